#!/bin/bash
# Round 2, GPU call 27 (the last two GPU-minutes): compute-sanitizer memcheck over the launch-bound kernels written this round --
# copy_resolved (bound and free entries), strip_resolved, box_gather_mail, flux_mail -- on a Bloch box and on the Nanoblock monitors.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( timeout 55 compute-sanitizer --tool memcheck --error-exitcode 9 tests/bin/monitors_mine_cuda nanoblock 12 > $O/r02_sanitize_monitors_nanoblock_v27.out 2> $O/r02_sanitize_monitors_nanoblock_v27.err; echo "exit $?" ) > $O/r02_sanitize_monitors_nanoblock_v27.log 2>&1
tail -3 $O/r02_sanitize_monitors_nanoblock_v27.out | cut -c1-200; cat $O/r02_sanitize_monitors_nanoblock_v27.log
( timeout 55 compute-sanitizer --tool memcheck --error-exitcode 9 tests/bin/monitors_mine_cuda phc 12 > $O/r02_sanitize_monitors_phc_v27.out 2> $O/r02_sanitize_monitors_phc_v27.err; echo "exit $?" ) > $O/r02_sanitize_monitors_phc_v27.log 2>&1
tail -3 $O/r02_sanitize_monitors_phc_v27.out | cut -c1-200; cat $O/r02_sanitize_monitors_phc_v27.log
