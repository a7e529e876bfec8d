#!/bin/bash
# Round-2 profiling pass on one B200 (run under gpurun): every command first runs plain and must exit 0, then under ncu.
# Outputs land in gpurun_out/; the summaries worth keeping are copied to profiles/ by hand.
set -u
O=gpurun_out
run() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > $O/${name}_plain.log 2>&1 || { echo "plain run of $name failed"; tail -5 $O/${name}_plain.log; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o $O/r02_${name} "$@" > $O/${name}_ncu.log 2>&1
  tail -1 $O/${name}_ncu.log
}
run rect_pair   "rectify_mono_pair"       3 python tools/kbench.py --only rect --batch 64 --iters 3 --one
run voxel       "voxel_cloud"             3 python tools/kbench.py --only voxel --batch 32 --iters 3 --one
run bp_colour   "backproject_vec_kernel"  6 python tools/kbench.py --only bp --batch 16 --iters 2
run reg_colour  "register_colour"         1 python tools/kbench.py --only bp --batch 16 --iters 2
run conv_nv12   "convert_vec_kernel"      10 python tools/kbench.py --only conv --batch 16 --iters 2
# launch list of the bench itself (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 5 --warmup 3 --no-rig --no-pcie --no-cpu > $O/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-rig --no-pcie --no-cpu > $O/bench_ncu.log 2>&1
tail -2 $O/bench_ncu.log
