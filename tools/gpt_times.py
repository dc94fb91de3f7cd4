"""In-situ timing of the GPT decode step (B utterances): wall time of the generate call, CUDA-event time of every launch of
one step, and the per-token cost of the step replayed as one CUDA graph -- for the fused step (csrc/gpt_dgemm.cu, with and
without programmatic dependent launch, several cluster-size choices) and the round-1 kernel-by-kernel step.
    B=16 python tools/gpt_times.py"""
import collections
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import detail_tts_b200.gpt as G  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
from detail_tts_b200.gpt import UnifiedVoice  # noqa: E402

B = int(os.environ.get("B", 128))
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
W = synth.synth_state_dict(0, keys=lambda k: k.startswith("gpt.") and synth.infer_path_key(k))
gpt = UnifiedVoice(W, dev)
text, refer = bench.make_inputs(B)
text, refer = text.to(dev), refer.to(dev)
kw = dict(do_sample=True, top_p=.8, temperature=.8, repetition_penalty=2.0, max_generate_length=71,
          text_lengths=[51] * B, suppress_tokens=[8193])


def run(label, fused, pdl=True, splits=None, detail=False):
    G.FUSED_STEP = fused
    if splits:
        G.FUSED_SPLITS = splits
    gpt.use_pdl_fused = pdl
    gpt._states.clear()
    walls = []
    for _ in range(3):
        torch.manual_seed(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gpt.inference_speech_tortoise(refer, [300] * B, text, **kw)
        torch.cuda.synchronize()
        walls.append(1000 * (time.perf_counter() - t0))
    st = list(gpt._states.values())[0]
    graph, plan = (st.loop_graph, st.loop_plan) if st.fused else (st.graph, st.plan)
    us = float("nan")
    if graph is not None:
        def rewind():
            st.step.zero_()
            st.unfinished.fill_(1)
        rewind()
        for _ in range(3):
            graph.replay()
        rewind()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1000
    print(f"[{label}] B={B}: generate wall {walls[0]:.1f} / {walls[1]:.1f} / {walls[2]:.1f} ms; step as one CUDA graph {us:.1f} us "
          f"({len(plan)} entry points)")
    if detail:
        st.step.zero_()
        plan.profile()
        st.step.zero_()
        res = plan.profile()
        st.step.zero_()
        agg = collections.defaultdict(lambda: [0, 0.0])
        for name, s, ms in res:
            key = name
            if "gemm" in name:
                key += f"_N{s.N}_K{s.K}"
            agg[key][0] += 1
            agg[key][1] += ms
        tot = sum(v[1] for v in agg.values())
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"   {v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:3d} avg {1000 * v[1] / v[0]:8.1f} us  {k}")
        print(f"   launch-by-launch total {tot:.3f} ms, {len(res)} launches")


default_splits = G.FUSED_SPLITS
run("fused + PDL", True, True, detail=True)
run("fused, no PDL", True, False)
for sp in [v for v in os.environ.get("SPLITS", "").split(";") if v]:
    run("fused + PDL splits " + sp, True, True, tuple(int(v) for v in sp.split(",")))
G.FUSED_SPLITS = default_splits
run("round-1 kernel-by-kernel", False, detail=True)
