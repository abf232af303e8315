set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "pc_steps or philox" tests/test_gpu_sampler.py -x -q > gpurun_out/ab5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab5_tests.log
tail -5 gpurun_out/ab5_tests.log
O=gpurun_out/ab5_steps.jsonl; : > $O
timeout 120 python tools/bench_steps.py --in-place --check build/ab/libt2p_old.so --tag inplace >> $O 2>&1
timeout 120 python tools/bench_steps.py --tag copy_through >> $O 2>&1
timeout 120 python tools/bench_steps.py --mask ones --tag ones >> $O 2>&1
timeout 120 python tools/bench_steps.py --in-place --B 128 --C 8 --tag inplace_cfg3 >> $O 2>&1
cat $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/ab5_fullgpu.log 2>&1; tail -5 gpurun_out/ab5_fullgpu.log
timeout 600 python bench.py > gpurun_out/ab5_bench.json 2> gpurun_out/ab5_bench.err; tail -c 3000 gpurun_out/ab5_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"predictor_kernel|corrector_kernel" --launch-skip 4 -c 2 -f -o gpurun_out/r01_pc_steps_v2 python tools/profile_steps.py > gpurun_out/ab5_ncu.log 2>&1
tail -3 gpurun_out/ab5_ncu.log
