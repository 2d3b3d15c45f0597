#!/bin/bash
# ncu --set full captures of the INT8 (tcgen05 kind::i8) and XNOR (popcount and +-1 kind::i8) convolution kernels:
# every convolution launch of one eager forward of the two BASELINE configs -> gpurun_out/*.ncu-rep + op lists.
set -e
cd "$(dirname "$0")/.."
export YB_NO_GRAPH=1
N="ncu --set full --clock-control none --import-source on -f"
python tools/run_forward.py --model yolov3-tiny --size 416 --batch 64 --quantized 1 --list > gpurun_out/oplist_tiny_int8.txt 2>&1
python tools/run_forward.py --model tiny-yolo-obj_xnor --size 416 --batch 64 --list > gpurun_out/oplist_xnor_tc.txt 2>&1
YB_XNOR_TC=0 python tools/run_forward.py --model tiny-yolo-obj_xnor --size 416 --batch 64 --list > gpurun_out/oplist_xnor_popc.txt 2>&1
$N -k regex:k_conv_tc -o gpurun_out/r01_int8_tiny416_b64_conv_tc python tools/run_forward.py --model yolov3-tiny --size 416 --batch 64 --quantized 1 > /dev/null 2>&1
$N -k regex:k_conv_tc -o gpurun_out/r01_xnor_tc_416_b64 python tools/run_forward.py --model tiny-yolo-obj_xnor --size 416 --batch 64 > /dev/null 2>&1
YB_XNOR_TC=0 $N -k regex:k_conv_xnor -o gpurun_out/r01_xnor_popc_416_b64 python tools/run_forward.py --model tiny-yolo-obj_xnor --size 416 --batch 64 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
