#!/bin/bash
# Round-end evidence run on one B200 (gpurun -- 'bash tools/final_profile.sh'): GPU tests, smoke, both bench arms, then
# -- each only after its command exited 0 without ncu -- launch lists and one full capture of the top kernel for the
# headline workload (c2) and for generation 0 of the ripa scene (c5). Outputs go to gpurun_out/ (tag $1, default r2f).
T=${1:-r2f}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${T}_gputest.log 2>&1; tail -2 $O/${T}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
timeout 900 python bench.py --impl reference > $O/${T}_bench_reference_arm.json 2> $O/${T}_bench_reference_arm.err; echo "ref arm rc=$?"
timeout 900 python bench.py > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err; echo "bench rc=$?"
C2="python bench.py --steps 2 --warmup 3 --only --e2e-steps 0 --flag-rays 0 --cpu-rays 1000"
if timeout 300 $C2 > $O/${T}_c2_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c2.csv $C2 > $O/${T}_c2_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o $O/${T}_c2 $C2 > $O/${T}_c2_ncu_full.log 2>&1
fi
C5="python tools/variant_bench.py --variants intree --workloads c5_ripa_64 --steps 2"
if timeout 300 $C5 > $O/${T}_c5_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${T}_launches_c5.csv $C5 > $O/${T}_c5_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 33 -c 1 -f -o $O/${T}_c5_gen0 $C5 > $O/${T}_c5_ncu_full.log 2>&1
fi
ls -la $O | grep ${T}_
