#!/bin/bash
# round-2 ncu evidence: launch list of the default bench command + ncu --set full of the kernels DESIGN.md discusses.
# The .ncu-rep files (80 MB together) are summarised ON the box (tools/ncu_summary.py) and removed: gpurun_out/ returns <= 64 MiB.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --nvtx --nvtx-include target/ -f"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_b32.csv python bench.py --steps 2 --warmup 3 --preheat 0 --cpu-frames 0 --plugin-frames 0 --extras "" > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
cap() {  # name, header, labels, then ncu_target args
  local name=$1 header=$2 labels=$3; shift 3
  timeout 900 $NCU -o gpurun_out/$name python tools/ncu_target.py "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.txt "$header" "$labels"
  rm -f gpurun_out/$name.ncu-rep
}
OPS_FAST="res.conv0,res.conv1+head,shuf8.conv+blur,layers.6.conv,middle.0,layers.0.6.9.conv1,layers.0.6.9.conv2,layers.0.6.9.conv3,layers.5.conv.3.logits,layers.5.conv.3.pv,pixel"
cap r02_ncu_cfg2_fast "ncu --set full --clock-control none: cfg2 (video rf 24, 1080p, B = 32, fp16, precision fast), one launch each (tools/ncu_target.py)" "$OPS_FAST,pre.v,post.v,post.h" --config cfg2 --precision fast --ops "$OPS_FAST"
OPS_BAL="stem.conv,layers.0.4.1.conv3,layers.0.6.9.conv1,layers.0.6.9.conv2,layers.0.6.9.conv3"
cap r02_ncu_cfg2_balanced "ncu --set full: cfg2, precision balanced = split-precision (hi+lo, 3 MMAs per K step) encoder launches" "$OPS_BAL" --config cfg2 --precision balanced --ops "$OPS_BAL"
OPS_DEEP="res.conv0,layers.7.conv1,layers.0.6.3.conv1,layers.5.conv2.3.pv"
cap r02_ncu_cfg5_deep "ncu --set full: cfg5 (UHD, artistic = DynamicUnetDeep @640, B = 4): launches of the deep generator" "$OPS_DEEP" --config cfg5 --batch 4 --prog prog2 --ops "$OPS_DEEP"
OPS_Z="model2.0,model5.0,model8.2"
cap r02_ncu_cfg5_zhang "ncu --set full: cfg5, Zhang eccv16 @256 (B = 4): split-precision early block, dilated block, tail" "$OPS_Z" --config cfg5 --batch 4 --prog zhang --ops "$OPS_Z"
du -sh gpurun_out
