#!/bin/bash
# Round-2 captures, run ON the GPU box (gpurun): usage  bash tools/capture_profiles_r2.sh <tag>
# Every ncu call is wrapped in `timeout`; .ncu-rep files are summarised and deleted (gpurun brings back <= 64 MiB).
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
for U in 16 128; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_raw_$U.csv \
      python tools/profile_step.py --utts $U > $OUT/${TAG}_ncu_launches_$U.log 2>&1
  timeout 60 python tools/ncu_summary.py launches $OUT/${TAG}_launches_raw_$U.csv $OUT/${TAG}_launches_summary_b$U.csv \
      "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off: python tools/profile_step.py --utts $U" \
      "slice: 2 diffusion sampler steps (2 batched cond+uncond evals) + GPT prefill + 3 decode steps (fused step for <= 64 utterances) + flow-VAE/vocoder, B=$U, T=70" \
      "per-launch times are cold-cache and serialised: compare SHARES, not absolutes"
  rm -f $OUT/${TAG}_launches_raw_$U.csv
done
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -c 62 --csv --log-file $OUT/${TAG}_gemm_metrics_raw.csv \
    python tools/profile_step.py --utts 128 --parts diffusion > $OUT/${TAG}_ncu_gemm_metrics.log 2>&1
timeout 60 python tools/ncu_summary.py metrics $OUT/${TAG}_gemm_metrics_raw.csv $OUT/${TAG}_gemm_metrics.csv \
    "ncu --metrics $M -k regex:gemm_tc_kernel -c 62: the 62 GEMM launches of one batched (cond+uncond) diffusion eval, B=128, F=280"
rm -f $OUT/${TAG}_gemm_metrics_raw.csv
cap() {  # name regex skip count parts utts
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 -o $OUT/${TAG}_$1 \
      python tools/profile_step.py --utts $6 --parts $5 > $OUT/${TAG}_ncu_$1.log 2>&1
  timeout 120 python tools/ncu_summary.py full $OUT/${TAG}_$1.ncu-rep $OUT/${TAG}_full_$1.csv \
      "ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4: python tools/profile_step.py --utts $6 --parts $5"
  rm -f $OUT/${TAG}_$1.ncu-rep
}
cap gpt_b16 "dgemm_kernel|attention_decode|process_logits|final_ln" 60 12 gpt 16
cap gpt_b64 "dgemm_kernel|attention_decode|process_logits|final_ln" 60 12 gpt 64
cap gemm "gemm_tc_kernel" 24 6 diffusion 128
cap attn_gn "flash48_tc|groupnorm" 10 4 diffusion 128
ls -la $OUT | tail -20
