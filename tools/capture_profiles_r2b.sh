#!/bin/bash
# Round-2 (second session) captures, run ON the GPU box (gpurun): usage  bash tools/capture_profiles_r2b.sh <tag>
# Reduced set (the GPU budget of the round was nearly spent): launch list of the 128-utterance slice, DRAM traffic / tensor-pipe
# metrics of the 62 GEMM launches of one batched diffusion eval, full captures of the GEMM and attention kernels.
# Every ncu call is wrapped in `timeout`; .ncu-rep files are summarised and deleted (gpurun brings back <= 64 MiB).
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
U=128
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_raw_$U.csv \
    python tools/profile_step.py --utts $U > $OUT/${TAG}_ncu_launches_$U.log 2>&1
timeout 60 python tools/ncu_summary.py launches $OUT/${TAG}_launches_raw_$U.csv $OUT/${TAG}_launches_summary_b$U.csv \
    "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off: python tools/profile_step.py --utts $U" \
    "slice: 2 diffusion sampler steps (2 batched cond+uncond evals) + GPT prefill + 3 decode steps + flow-VAE/vocoder, B=$U, T=70" \
    "per-launch times are cold-cache and serialised: compare SHARES, not absolutes"
rm -f $OUT/${TAG}_launches_raw_$U.csv
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -c 62 --csv --log-file $OUT/${TAG}_gemm_metrics_raw.csv \
    python tools/profile_step.py --utts 128 --parts diffusion > $OUT/${TAG}_ncu_gemm_metrics.log 2>&1
timeout 60 python tools/ncu_summary.py metrics $OUT/${TAG}_gemm_metrics_raw.csv $OUT/${TAG}_gemm_metrics.csv \
    "ncu --metrics $M -k regex:gemm_tc_kernel -c 62: the 62 GEMM launches of one batched (cond+uncond) diffusion eval, B=128, F=280"
rm -f $OUT/${TAG}_gemm_metrics_raw.csv
cap() {  # name regex skip count parts utts
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 -o $OUT/${TAG}_$1 \
      python tools/profile_step.py --utts $6 --parts $5 > $OUT/${TAG}_ncu_$1.log 2>&1
  timeout 120 python tools/ncu_summary.py full $OUT/${TAG}_$1.ncu-rep $OUT/${TAG}_full_$1.csv \
      "ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4: python tools/profile_step.py --utts $6 --parts $5"
  rm -f $OUT/${TAG}_$1.ncu-rep
}
cap gemm "gemm_tc_kernel" 24 6 diffusion 128
cap attn "flash48_tc" 10 1 diffusion 128
ls -la $OUT | tail -12
