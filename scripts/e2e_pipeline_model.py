"""Model of the host-mode pipeline of rfb_lu_f64 on a page-locked matrix (no GPU needed): the compute stream executes
the host driver's own schedule (rfb_trace_lu) with MEASURED per-class kernel costs (round-2 launch list / bench
profile), the download stream sends every early-download tile at the measured PCIe rate, in issue order, as soon as
the compute stream has reached the point where the driver enqueues it.  Prints when the last panel ends and when the
last byte has arrived for the two early-download schemes (1 = row bands at the right spine, 2 = finished tiles) --
the difference is what the scheme costs on top of the device-resident time -- and how long the compute stream waits for
upload chunks at the start.  A ranking tool, not a predictor."""
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rfb200  # noqa: E402

PANEL, LASWP, TRSM, GEMM, DOWNLOAD = 1, 3, 4, 5, 6
rate_by_k = {64: 15.8, 128: 22.0, 256: 26.9, 512: 30.1, 1024: 31.7, 2048: 32.7, 4096: 34.0, 8192: 35.2}   # TFLOP/s, measured


def gemm_s(m, nn, k):
    if m <= 0 or nn <= 0 or k <= 0:
        return 0.0
    kk = min(rate_by_k, key=lambda x: abs(np.log2(x) - np.log2(max(k, 64))))
    tiles = -(-m // 128) * -(-nn // 128)
    waves = -(-tiles // 148)
    fill = tiles / (waves * 148.0)
    width = min(1.0, nn / (128.0 * -(-nn // 128)))
    return 2.0 * m * nn * k / (rate_by_k[kk] * min(1.0, fill / 0.9) * width * 1e12) + 4e-6


def trsm_s(k, nrhs):
    """host-recursive blocking at multiples of 256 (csrc/trsm.cu): diagonal blocks 30.5 us per 256 rows (12.7 / 6.4 us
    for 128 / 64), off-diagonal work as GEMMs"""
    if k <= 64:
        return 6.4e-6
    if k <= 128:
        return 12.7e-6
    if k <= 256:
        return 30.5e-6 * (1.0 if nrhs <= 4096 else 1.45)
    k1 = (k // 2 + 255) // 256 * 256
    return trsm_s(k1, nrhs) + gemm_s(k - k1, nrhs, k1) + trsm_s(k - k1, nrhs)


def upload_schedule(n, s, pcie_up, chunk_overhead=12e-6):
    """column bounds and completion times of the upload chunks (csrc/rfb_api.cu: 16 chunks of 8 MB, then doubling to 64 MB)"""
    col_bytes, target, j, bounds = s * n, 8 << 20, 0, []
    while j < n:
        j = min(n, j + max(64, target // col_bytes))
        bounds.append(j)
        if len(bounds) >= 16 and target < (64 << 20):
            target *= 2
    t, prev, done = 0.0, 0, []
    for b in bounds:
        t += chunk_overhead + (b - prev) * col_bytes / pcie_up
        done.append(t)
        prev = b
    return bounds, done


def simulate(n, mode, s=8, pcie_down=53e9, copy_overhead=8e-6, pcie_up=55e9, upload=False, stats=None):
    ops = rfb200.trace_lu(n, n, pinned_host=True, early_mode=mode)
    bounds, done = upload_schedule(n, s, pcie_up)

    def resident(col):            # when columns [0, col) have arrived
        for b, d in zip(bounds, done):
            if b >= col:
                return d
        return done[-1]

    t = 0.0                       # compute stream clock
    down_free = 0.0               # download stream clock
    last_panel_end = 0.0
    bytes_down = 0
    stalled = 0.0
    for op, r, c, s0, s1, s2, r2, c2 in ops.tolist():
        if upload and op in (PANEL, LASWP, TRSM, GEMM):       # the compute stream waits for the columns the kernel touches
            need = {PANEL: c + s1, LASWP: c + s0, TRSM: c2 + s1, GEMM: c + s1}[op]
            u = resident(need)
            if u > t:
                stalled += u - t
                t = u
        if op == PANEL:
            t += s1 * 1.9e-6 + 8e-6
            last_panel_end = t
        elif op == LASWP:
            p = s2 - s1
            t += 5e-6 + (17e-6 if p >= 512 else 0.0) + 4.0 * s * p * s0 / 3.0e12
        elif op == TRSM:
            t += trsm_s(s0, s1)
        elif op == GEMM:
            t += gemm_s(s0, s1, s2)
        elif op == DOWNLOAD:
            nbytes = s0 * s1 * s
            eff = min(1.0, (s0 * s) / 4096.0) ** 0.5        # narrow 2-D segments lose DMA efficiency (assumed, not measured)
            down_free = max(down_free, t) + copy_overhead + nbytes / (pcie_down * eff)
            bytes_down += nbytes
    if stats is not None:
        stats["stalled_on_uploads_ms"] = 1e3 * stalled
        stats["upload_done_ms"] = 1e3 * done[-1]
    end_compute = t
    tail_bytes = n * n * s - bytes_down
    end = max(end_compute, down_free) + tail_bytes / pcie_down
    return end_compute, last_panel_end, end, tail_bytes


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    for mode in (0, 1, 2):
        ec, lp, end, tail = simulate(n, mode)
        print(f"n = {n} early-download mode {mode}: compute stream ends at {ec * 1e3:7.2f} ms, last byte on the host at "
              f"{end * 1e3:7.2f} ms (+{(end - ec) * 1e3:5.2f} ms), {tail / 1e6:7.1f} MB left for the final copy")
    st = {}
    ec_up, _, _, _ = simulate(n, 2, upload=True, stats=st)
    print(f"n = {n} with the pipelined upload at 55 GB/s: the compute stream waits {st['stalled_on_uploads_ms']:.2f} ms for columns that "
          f"have not arrived yet (upload complete at {st['upload_done_ms']:.1f} ms): at the start a 64-column step costs less than "
          f"its columns take to upload")
