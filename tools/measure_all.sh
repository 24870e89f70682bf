#!/bin/bash
# one GPU call that regenerates the measured evidence of a round: usage  bash tools/measure_all.sh r3
# (writes gpurun_out/<tag>_*; the summaries that go under profiles/ are produced from these by tools/ncu_summary.py and by hand)
tag=${1:-rX}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gpu_pytest.txt 2>&1; tail -2 gpurun_out/${tag}_gpu_pytest.txt
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench_n1.err
timeout 600 python tools/bench_all.py > gpurun_out/${tag}_all_primitives_10bit.md 2> gpurun_out/${tag}_all10.err
timeout 600 python tools/bench_all.py --depth 8 > gpurun_out/${tag}_all_primitives_8bit.md 2> gpurun_out/${tag}_all8.err
timeout 600 python tools/bench_all.py --depth 12 > gpurun_out/${tag}_all_primitives_12bit.md 2> gpurun_out/${tag}_all12.err
timeout 600 python tools/bench_me.py --methods 5,1,0,2,3,4 --shapes 64,16,8 --bidir --lowres --cpu > gpurun_out/${tag}_motion_search_10bit.json 2> gpurun_out/${tag}_me10.err
timeout 300 python tools/bench_tu.py 10 16 > gpurun_out/${tag}_tu_chain.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --traffic off > gpurun_out/${tag}_launches_bench.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/${tag}_smi.txt
echo done
