timeout 600 python -m pytest tests/test_gpu_batch2.py -x -q -k "tu_chain" > gpurun_out/t_tu.log 2>&1; tail -4 gpurun_out/t_tu.log
timeout 300 python tools/bench_all.py --only tu > gpurun_out/b_tu.md 2> gpurun_out/b_tu.err; grep tu_chain gpurun_out/b_tu.md
for k in 0 1 2 3 4; do echo "nquant knob $k"; X265B200_LAB=$k timeout 200 python tools/bench_all.py --only quant 2>/dev/null | grep nquant; done
for k in "0,0,0" "0,1,0" "0,0,2" "0,0,4" "0,1,2" "0,1,4" "0,0,1"; do echo "sa8d knob $k"; X265B200_LAB=$k timeout 200 python tools/bench_all.py --only metrics 2>/dev/null | grep sa8d; done
