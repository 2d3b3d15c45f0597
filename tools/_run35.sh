for w in yolov3-tiny-416-int8-b64 tiny-yolo-obj_xnor-416-b64; do
  for b in 1 2; do
    YB_TC_EPI_BUFS=$b timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 30 > gpurun_out/r35_${w}_b$b.json 2> gpurun_out/r35_${w}_b$b.err
    echo "$w bufs $b rc=$?"; tail -c 600 gpurun_out/r35_${w}_b$b.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('no json', e)"
    tail -5 gpurun_out/r35_${w}_b$b.err
  done
done
