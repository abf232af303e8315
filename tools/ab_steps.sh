set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "pc_steps or philox" tests/test_gpu_sampler.py -x -q > gpurun_out/ab1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab1_tests.log
tail -5 gpurun_out/ab1_tests.log
O=gpurun_out/ab1_steps.jsonl; : > $O
timeout 120 python tools/bench_steps.py --lib build/ab/libt2p_old.so --tag old >> $O 2>&1
timeout 120 python tools/bench_steps.py --check build/ab/libt2p_old.so --tag new >> $O 2>&1
T2P_STEP_NOSKIP=1 timeout 120 python tools/bench_steps.py --tag noskip >> $O 2>&1
T2P_STEP_NOCACHE=1 timeout 120 python tools/bench_steps.py --tag nocache >> $O 2>&1
T2P_STEP_NOPREFETCH=1 timeout 120 python tools/bench_steps.py --tag noprefetch >> $O 2>&1
timeout 120 python tools/bench_steps.py --mask ones --tag new_ones >> $O 2>&1
timeout 120 python tools/bench_steps.py --lib build/ab/libt2p_old.so --mask ones --tag old_ones >> $O 2>&1
timeout 120 python tools/bench_steps.py --B 128 --C 8 --tag new_cfg3 >> $O 2>&1
timeout 120 python tools/bench_steps.py --lib build/ab/libt2p_old.so --B 128 --C 8 --tag old_cfg3 >> $O 2>&1
cat $O
