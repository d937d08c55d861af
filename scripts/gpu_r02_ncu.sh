#!/bin/bash
# round 2 ncu evidence: full-set captures of the Net A kernels and of Net B's bf16 layers, launch list of the bench command
mkdir -p gpurun_out
# barrier probes one row ahead (default) vs blocking waits at the head of every row (tc_debug=8)
for dbg in 0 8 4 12; do timeout 120 python scripts/prof_tc.py --batch 256 --opt tc_debug=$dbg 2>&1 | tail -2; done > gpurun_out/r02_probe_ab.txt 2>&1
cat gpurun_out/r02_probe_ab.txt | cut -c1-400
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_rollout_gpu.py tests/test_estimator_gpu.py tests/test_latband_gpu.py -q -x 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_sw_kernel -s 2 -c 2 -f -o gpurun_out/r02_prof_net_a python scripts/prof_tc.py --batch 256 --iters 1 2>&1 | tail -1
DLWP_PRECISION=bf16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_sw_kernel|p_ew_kernel" -s 12 -c 12 -f -o gpurun_out/r02_prof_net_b_bf16 python scripts/bench_net_b.py --batch 64 --steps 2 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r02_ncu_bench.log
timeout 300 python bench.py --steps 200 --warmup 3 > gpurun_out/r02_bench200.log 2>&1; tail -1 gpurun_out/r02_bench200.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench20.log 2>&1; tail -1 gpurun_out/r02_bench20.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.log 2>&1; tail -1 gpurun_out/r02_bench_reference.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
