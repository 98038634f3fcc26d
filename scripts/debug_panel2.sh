for mode in 0 1 2 3; do echo "== RFB_XMODE=$mode"; RFB_XMODE=$mode timeout 120 python scripts/debug_panel.py 2>&1 | tail -5 | cut -c1-150; done
