#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun): one `--set full` capture per headline kernel, summarised into
# gpurun_out/r2_<name>_ncu.txt by tools/ncu_summary.py; the .ncu-rep is kept only for the names listed in KEEP.
#   bash tools/r2_ncu.sh [name ...]      (default: all)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KEEP="roi_align resize128 deform_bf16"
declare -A OP=( [roi_align]=roi_align [resize128]=resize128 [deform_bf16]=deform [deform_f32]=deform_f32 [bnms_mask]=batched_nms
                [bnms_scan]=batched_nms [roi_pool]=roi_pool [ps_roi_align]=ps_roi_align [roi_align_bwd]=roi_align_bwd
                [roi_align_bwd_det]=roi_align_bwd_det [multiscale]=multiscale [preprocess]=preprocess )
declare -A KER=( [roi_align]=roi_align_line_kernel [resize128]=resize_aa_stream_kernel [deform_bf16]=deform_conv2d_tc_kernel
                 [deform_f32]=deform_conv2d_tc3_kernel [bnms_mask]=bnms_mask_kernel [bnms_scan]=bnms_scan_kernel
                 [roi_pool]=roi_pool_plane_kernel [ps_roi_align]=ps_roi_align_plane_kernel [roi_align_bwd]=roi_align_bwd_plane_atomic_kernel
                 [roi_align_bwd_det]=roi_align_bwd_plane_fast_kernel [multiscale]=roi_align_line_kernel [preprocess]=resize_crop_norm_kernel )
NAMES="${@:-roi_align resize128 deform_bf16 deform_f32 bnms_mask bnms_scan roi_pool ps_roi_align roi_align_bwd roi_align_bwd_det multiscale preprocess}"
for n in $NAMES; do
  rm -f gpurun_out/r2_$n.ncu-rep
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KER[$n]} --launch-skip 2 -c 1 -f -o gpurun_out/r2_$n \
      python tools/prof_ops.py ${OP[$n]} 1 > gpurun_out/r2_${n}_ncu.log 2>&1
  if [ -f gpurun_out/r2_$n.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/r2_$n.ncu-rep > gpurun_out/r2_${n}_ncu.txt 2>&1
    case " $KEEP " in *" $n "*) ;; *) rm -f gpurun_out/r2_$n.ncu-rep ;; esac
  else
    echo "no report for $n" > gpurun_out/r2_${n}_ncu.txt; tail -5 gpurun_out/r2_${n}_ncu.log >> gpurun_out/r2_${n}_ncu.txt
  fi
done
ls -la gpurun_out/r2_*_ncu.txt
