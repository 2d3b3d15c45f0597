timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -4
for v in 0 1 2; do YB_TC_CG2_BN128=$v python tools/run_forward.py --list --reps 3 > gpurun_out/r33_cg$v.txt 2>&1; done
echo "--- cg2_bn128 0 vs 1"; python tools/ab_layers.py yolov3 608 gpurun_out/r33_cg0.txt gpurun_out/r33_cg1.txt | grep -E "<--|total"
echo "--- cg2_bn128 0 vs 2"; python tools/ab_layers.py yolov3 608 gpurun_out/r33_cg0.txt gpurun_out/r33_cg2.txt | grep -E "<--|total"
STEPS=30 BENCH_ARGS=--no-cpu-baseline bash tools/ab_bench.sh "cg0:YB_TC_CG2_BN128=0" "cg1:YB_TC_CG2_BN128=1" "cg2:YB_TC_CG2_BN128=2" "cg0:YB_TC_CG2_BN128=0" "cg2:YB_TC_CG2_BN128=2"
YB_TC_STATS=1 python tools/run_forward.py 2>&1 | grep "TCSTATS" | head -12
