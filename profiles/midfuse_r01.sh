timeout 200 python profiles/small_grid_probe.py 256 320 384 512 640 2>&1 | tail -5
ALISB_FUSE=0 timeout 200 python profiles/small_grid_probe.py 256 320 384 512 640 2>&1 | tail -5
