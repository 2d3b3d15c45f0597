for v in "YB_TC_NO_S2_TMA_EPI=1" "YB_TC_S2_CG1=1" "YB_TC_S2_HALO_MAXC=0" "YB_TC_S2_HALO_MAXC=0 YB_TC_NO_S2_TMA_EPI=1"; do
  echo "=== $v"; env $v timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "every_layer or stride2" 2>&1 | tail -4
done
echo "=== sanitizer"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "every_layer and 64-2" 2>&1 | grep -v "^$" | head -40
