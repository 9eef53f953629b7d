set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r1c_pytest.log 2>&1; tail -3 gpurun_out/r1c_pytest.log
timeout 300 python __graft_entry__.py > gpurun_out/r1c_smoke.log 2>&1; tail -2 gpurun_out/r1c_smoke.log
timeout 600 python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; tail -c 600 gpurun_out/r1c_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1c_bench_ref.json 2>> gpurun_out/r1c_bench.err; cat gpurun_out/r1c_bench_ref.json | tail -c 400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r1c_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_tc -s 3 -c 1 -o gpurun_out/prof_r1c_decoder -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r1c_ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lift_kernel|sample_staged|nchw_to_nhwc|fuse_kernel" -c 8 -o gpurun_out/prof_r1c_lift_sample -f python tools/prof_kernels.py all 1 > gpurun_out/r1c_ncu_s.log 2>&1
ls -la gpurun_out | tail -12
