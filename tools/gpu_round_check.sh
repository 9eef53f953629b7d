# Round-end check on the GPU box: tests, smoke, both bench arms, launch list and ncu captures (names: r1e_*).
set -x
T=r1g
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 1500 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench_ref.json | tail -c 400
timeout 300 python tools/sample_bench.py binned staged > gpurun_out/${T}_sample_bench.txt 2>&1; cat gpurun_out/${T}_sample_bench.txt
timeout 300 python tools/scatter_bench.py > gpurun_out/${T}_scatter_bench.txt 2>&1; cat gpurun_out/${T}_scatter_bench.txt
timeout 300 python tools/fusion_bench.py > gpurun_out/${T}_fusion_bench.txt 2>&1; tail -3 gpurun_out/${T}_fusion_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_tc -s 3 -c 1 -o gpurun_out/prof_${T}_decoder -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lift_kernel|sample_staged|nchw_to_nhwc" -c 6 -o gpurun_out/prof_${T}_lift_sample -f python tools/prof_kernels.py all 1 > gpurun_out/${T}_ncu_s.log 2>&1
SAMPLE_CASES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sample_binned|bin_count|bin_scatter|bin_reduce|bin_scan" -s 5 -c 5 -o gpurun_out/prof_${T}_binned -f python tools/sample_bench.py child binned > gpurun_out/${T}_ncu_b.log 2>&1
ls -la gpurun_out | tail -14
