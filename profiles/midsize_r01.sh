timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chi2" 2>&1 | tail -3
for n in 256 512; do
ALISB_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cube${n}.csv python profiles/midsize_probe.py $n 3 > /dev/null 2>&1
done
python profiles/small_grid_probe.py 2>&1 | tail -8
ALISB_GRAPH=0 python profiles/small_grid_probe.py 2>&1 | tail -8
ALISB_CONC=0 python profiles/small_grid_probe.py 2>&1 | tail -8
