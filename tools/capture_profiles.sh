#!/bin/bash
# Run ON the GPU box (gpurun): ncu launch list + per-kernel captures, summarised to small CSVs under gpurun_out/
# (the .ncu-rep files are deleted: gpurun only brings back 64 MiB).  usage: bash tools/capture_profiles.sh <tag>
TAG=${1:-cap}
OUT=gpurun_out
mkdir -p $OUT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
# 1. every launch of a representative slice with its device time
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_raw.csv \
    python tools/profile_step.py --utts 64 > $OUT/${TAG}_ncu1.log 2>&1
python tools/ncu_summary.py launches $OUT/${TAG}_launches_raw.csv $OUT/${TAG}_launches_summary.csv \
    "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off: python tools/profile_step.py --utts 64" \
    "slice: 2 diffusion sampler steps (2 batched cond+uncond evals) + GPT prefill + 3 decode steps + flow-VAE/vocoder, B=64, T=70" \
    "per-launch times are cold-cache and serialised: compare SHARES, not absolutes"
rm -f $OUT/${TAG}_launches_raw.csv
# 2. DRAM traffic + tensor-pipe activity of every tcgen05 GEMM launch of one batched diffusion eval at the bench shape
ncu --metrics $M --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -c 62 --csv --log-file $OUT/${TAG}_gemm_metrics_raw.csv \
    python tools/profile_step.py --utts 128 --parts diffusion > $OUT/${TAG}_ncu2.log 2>&1
python tools/ncu_summary.py metrics $OUT/${TAG}_gemm_metrics_raw.csv $OUT/${TAG}_gemm_metrics.csv \
    "ncu --metrics $M -k regex:gemm_tc_kernel -c 62: the 62 GEMM launches of one batched (cond+uncond) diffusion eval, B=128, F=280"
rm -f $OUT/${TAG}_gemm_metrics_raw.csv
# 3. full-set captures of the hot kernels (a few launches each)
cap() {  # name regex skip count parts utts
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 -o $OUT/${TAG}_$1 \
      python tools/profile_step.py --utts $6 --parts $5 > $OUT/${TAG}_ncu_$1.log 2>&1
  python tools/ncu_summary.py full $OUT/${TAG}_$1.ncu-rep $OUT/${TAG}_full_$1.csv \
      "ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4: python tools/profile_step.py --utts $6 --parts $5"
  rm -f $OUT/${TAG}_$1.ncu-rep
}
cap gemm "gemm_tc_kernel" 24 6 diffusion 128
cap attn_gn "flash48_tc|groupnorm|pstep" 10 5 diffusion 128
cap voc "voc_mrf|conv_post" 0 3 vocoder 128
cap gpt "attention_decode|reduce_kernel|process_logits" 40 6 gpt 128
ls -la $OUT
