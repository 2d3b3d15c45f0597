timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
python tools/run_forward.py --list --reps 3 > gpurun_out/r34_b2.txt 2>&1
YB_TC_EPI_BUFS=1 python tools/run_forward.py --list --reps 3 > gpurun_out/r34_b1.txt 2>&1
YB_TC_EPI_BUFS=2 python tools/run_forward.py --list --reps 3 > gpurun_out/r34_b2all.txt 2>&1
echo "--- bufs 1 vs default(2 for SW32)"; python tools/ab_layers.py yolov3 608 gpurun_out/r34_b1.txt gpurun_out/r34_b2.txt | grep -E "<--|total"
echo "--- bufs 1 vs 2 everywhere"; python tools/ab_layers.py yolov3 608 gpurun_out/r34_b1.txt gpurun_out/r34_b2all.txt | grep -E "<--|total"
STEPS=30 BENCH_ARGS=--no-cpu-baseline bash tools/ab_bench.sh "b1:YB_TC_EPI_BUFS=1" "default:" "b2all:YB_TC_EPI_BUFS=2" "b1:YB_TC_EPI_BUFS=1" "default:"
for w in yolov3-tiny-416-int8-b64 tiny-yolo-obj_xnor-416-b64; do
STEPS=30 BENCH_ARGS="--no-cpu-baseline --workload $w" bash tools/ab_bench.sh "b1:YB_TC_EPI_BUFS=1" "b2:YB_TC_EPI_BUFS=2" "b1:YB_TC_EPI_BUFS=1" "b2:YB_TC_EPI_BUFS=2"
done
