#!/bin/bash
# ncu evidence for one round: launch list of one eager agent step + `--set full` captures of the dominant kernels.
# Usage (GPU box): bash tools/profile_round.sh r1g    -> gpurun_out/<tag>_*.txt (+ .ncu-rep)
TAG=${1:-r1x}
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/one_step.py > gpurun_out/${TAG}_one_step.log 2>&1
python tools/agg_launches.py gpurun_out/launches_${TAG}.csv --grids > gpurun_out/${TAG}_launches.txt 2>&1
# DRAM bytes per kernel of the same step (source of roofline.traffic in bench.py)
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/dram_${TAG}.csv python tools/one_step.py > gpurun_out/${TAG}_dram_one_step.log 2>&1
python tools/agg_dram.py gpurun_out/dram_${TAG}.csv > gpurun_out/${TAG}_dram_traffic.json 2>&1
rm -f gpurun_out/dram_${TAG}.csv
cap() {  # name, title, args...
  local name=$1 title=$2; shift 2
  ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/ncu_${TAG}_${name} \
      python tools/one_gemm.py "$@" > gpurun_out/${TAG}_${name}.log 2>&1
  python tools/ncu_summary.py gpurun_out/ncu_${TAG}_${name}.ncu-rep "$title" > gpurun_out/${TAG}_ncu_${name}.txt 2>&1
  # gpurun brings back at most 64 MiB: keep the reports of the two dominant kernels only, the summaries of all
  case "$name" in conv64|attn) ;; *) rm -f gpurun_out/ncu_${TAG}_${name}.ncu-rep ;; esac
}
cap conv64 "gn_conv2d 1x64x64 320->320 3x3 (U-Net level 0 ResBlock conv; split-K flavour)" conv 1 64 320 320
cap conv16 "gn_conv2d 1x16x16 1280->1280 3x3 (U-Net level 2 ResBlock conv; weight-streaming, split-K flavour)" conv 1 16 1280 1280
cap conv512 "gn_conv2d 1x512x512 256->256 3x3 (largest single op of the step, VAE up-block)" conv 1 512 256 256
cap qkv "gn_linear 4096x960x320 (U-Net level 0 fused QKV projection; compact flavour, TMA-stored tile)" linear 4096 960 320 plain
cap geglu "gn_linear 4096x2560x320 GEGLU (U-Net level 0 feed-forward)" linear 4096 2560 320 geglu
cap attn "gn_attention Tq=Tk=4096 heads=5 (U-Net level 0 self-attention)" attn 4096 4096 5
cap gnapply "gn_group_norm_apply 64x64x320 (+SiLU) from epilogue statistics" gnapply 64 320 10
cap conv8 "gn_conv2d 1x8x8 1280->1280 3x3 (U-Net level 3 / mid ResBlock conv; 29.5 MB of weights, split-K flavour)" conv 1 8 1280 1280
cap conv128 "gn_conv2d 1x128x128 512->512 3x3 (VAE up-block; CTA-pair flavour where the tile search picks it)" conv 1 128 512 512
cap ff2 "gn_linear 4096x320x1280 + bias + residual (U-Net level 0 feed-forward output projection)" linear 4096 320 1280 bias_res
cap xattn "gn_attention_qproj Tq=4096 Tk=77 heads=5 (U-Net level 0 cross-attention, query projection inside)" xattn 4096 77 5
cap attn1k "gn_attention Tq=Tk=1024 heads=10 (U-Net level 1 self-attention)" attn 1024 1024 10
cap ln "gn_layer_norm 258x512 (ACT transformer)" ln 258 512
cap softmax "gn_softmax_rows 4096x4096 fp32 -> fp16 (VAE mid-block attention)" softmax 4096 4096
cap u8 "u8_to_nhwc / nhwc_to_u8 512x512 (VaeImageProcessor pre / post-processing)" u8
cap tile "tile_views / untile_views 4 x 256x256 (controller/utils/misc.py on the device)" tile
ls -la gpurun_out | grep ${TAG}
