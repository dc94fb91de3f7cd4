import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from detail_tts_b200 import synth
from detail_tts_b200.model import SynthesizerTrn
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
B = 128
model = SynthesizerTrn(synth.synth_state_dict(0, keys=synth.infer_path_key), device=dev)
text, refer = bench.make_inputs(B)
tl, rl = [51] * B, [300] * B
kw = dict(max_generate_length=71, suppress_eos=True, do_sample=True)
text_d, refer_d = text.to(dev), refer.to(dev)
for it in range(8):
    s0 = torch.cuda.memory_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr = {"timing": True}
    model.infer_batch(text_d, tl, refer_d, rl, trace=tr, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    d = {k: s1[k] - s0[k] for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    print(f"wall {1000*(t1-t0):7.1f} ms", {k: round(v) for k, v in tr["stage_ms"].items()}, d,
          f"reserved {s1['reserved_bytes.all.current']/2**30:.1f} GiB peak_alloc {s1['allocated_bytes.all.peak']/2**30:.1f} GiB")
