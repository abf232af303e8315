set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "pc_steps or philox" tests/test_gpu_sampler.py -x -q > gpurun_out/ab6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab6_tests.log
tail -5 gpurun_out/ab6_tests.log
O=gpurun_out/ab6_steps.jsonl; : > $O
timeout 120 python tools/bench_steps.py --in-place --tag inplace_prefetch >> $O 2>&1
T2P_STEP_NOPREFETCH=1 timeout 120 python tools/bench_steps.py --in-place --tag inplace_noprefetch >> $O 2>&1
timeout 120 python tools/bench_steps.py --mask ones --tag ones_prefetch >> $O 2>&1
T2P_STEP_NOPREFETCH=1 timeout 120 python tools/bench_steps.py --mask ones --tag ones_noprefetch >> $O 2>&1
timeout 120 python tools/bench_steps.py --in-place --B 128 --C 8 --tag inplace_cfg3 >> $O 2>&1
timeout 120 python tools/bench_steps.py --in-place --B 8 --tag inplace_b8 >> $O 2>&1
cat $O
