#!/bin/bash
# Round 2, GPU call 1: whole GPU suite (incl. the reference's own 15 drivers), smoke, both bench arms, launch list,
# ADE-fused vs separate-kernel A/B on the 512^3 + Au block variant, ncu captures of the new kernels, sanitizer runs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r02_v1_gpu.txt 2>&1
nproc >> $O/r02_v1_gpu.txt; free -g | head -2 >> $O/r02_v1_gpu.txt
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 ) > $O/r02_pytest_gpu_v1.log 2>&1
tail -30 $O/r02_pytest_gpu_v1.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke_v1.log 2>&1; tail -2 $O/r02_smoke_v1.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1_v1.json 2> $O/r02_bench_n1_v1.err; tail -c 1500 $O/r02_bench_n1_v1.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref_v1.json 2> $O/r02_bench_ref_v1.err; tail -c 600 $O/r02_bench_ref_v1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_n1_v1.csv \
	python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_launches_bench_n1_v1.out 2>&1
# 512^3 vacuum+CPML cube, then the same with a gold block on 1/8 of the interior: ADE inside the volume kernels vs the three box kernels
B=tests/bin/bench_cube_mine_cuda
for v in "" "au"; do for fuse in 1 0; do
	[ -z "$v" ] && [ $fuse = 0 ] && continue
	echo "variant=[$v] ALISB_ADE_FUSE=$fuse $(ALISB_ADE_FUSE=$fuse $B 512 512 512 40 5 0 $v | tail -1)" >> $O/r02_ade_bench_v1.txt
done; done
cat $O/r02_ade_bench_v1.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02_launches_au512_v1.csv $B 512 512 512 3 1 0 au > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 8 --launch-count 4 -k regex:"yee_volume" -o $O/r02_ncu_ade_full $B 512 512 512 3 1 0 au > $O/r02_ncu_ade.log 2>&1
ncu -i $O/r02_ncu_ade_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active > $O/r02_ncu_ade_raw.csv 2>&1
# the phasor kernel (Dipole_Band: 81 frequencies, three faces)
mkdir -p /tmp/db && cp tests/bin/refex_Dipole_Band_cuda /tmp/db/Dipole_Band
( cd /tmp/db && timeout 300 ncu --set full --clock-control none --launch-skip 50 --launch-count 2 -k regex:"phaser_accumulate" -o $OLDPWD/$O/r02_ncu_phaser_full ./Dipole_Band > $OLDPWD/$O/r02_ncu_phaser.log 2>&1 )
ncu -i $O/r02_ncu_phaser_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size > $O/r02_ncu_phaser_raw.csv 2>&1
tail -3 $O/r02_ncu_phaser_raw.csv
# sanitizer: memcheck + racecheck over the volume / fused / list / phasor / monitor kernels on small grids
for tool in memcheck racecheck; do
	timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest \
		"tests/test_gpu_parity.py::test_box_world_seeded_bit_exact" "tests/test_gpu_parity.py::test_phaser_accumulation_bit_exact" \
		tests/test_monitor_eval.py tests/test_farfield_sums.py -m gpu -q -x > $O/r02_sanitize_${tool}_v1.log 2>&1
	echo "$tool rc=$?" >> $O/r02_sanitize_${tool}_v1.log; tail -4 $O/r02_sanitize_${tool}_v1.log
done
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitize_smoke_memcheck_v1.log 2>&1; echo "rc=$?" >> $O/r02_sanitize_smoke_memcheck_v1.log; tail -3 $O/r02_sanitize_smoke_memcheck_v1.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitize_smoke_racecheck_v1.log 2>&1; echo "rc=$?" >> $O/r02_sanitize_smoke_racecheck_v1.log; tail -3 $O/r02_sanitize_smoke_racecheck_v1.log
mkdir -p /tmp/nb && cp tests/bin/stepdump_mine_cuda /tmp/nb/ && ( cd /tmp/nb && timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 ./stepdump_mine_cuda nanoblock 40 /tmp/nb/out.bin > $OLDPWD/$O/r02_sanitize_nanoblock_memcheck_v1.log 2>&1; echo "rc=$?" >> $OLDPWD/$O/r02_sanitize_nanoblock_memcheck_v1.log )
tail -3 $O/r02_sanitize_nanoblock_memcheck_v1.log
