import json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from detail_tts_b200 import synth, dist as ddist
from detail_tts_b200.model import SynthesizerTrn, SynthPipeline
from detail_tts_b200.text import pad_ids
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
model = SynthesizerTrn(synth.synth_state_dict(0, keys=synth.infer_path_key), device=dev)
items = json.load(open("tests/golden/cfg4_mixed.json"))["items"]
text, tl = pad_ids([it["ids"] for it in items])
_, refer = bench.make_inputs(128)
nb = int(os.environ.get("NB", 2))
mode = os.environ.get("MODE", "pipe")
shards = ddist.shard_slices(128, nb, costs=tl)
mine = shards[0]
idx = torch.tensor(mine)
t, rf = text[idx].to(dev), refer[idx].to(dev)
tlm = [tl[i] for i in mine]
kw = dict(max_generate_length=71, suppress_eos=True, do_sample=True)
pipe = SynthPipeline(model)
if os.environ.get("EQUAL"):
    t, tlm = t[:, :52].contiguous(), [51] * len(mine)
for rep in range(int(os.environ.get("REPS", 5))):
    torch.manual_seed(1)
    if mode == "overlap":
        for k in range(6):
            wav, wl = pipe.submit(t, tlm, rf, [300] * len(mine), **kw)
        pipe.drain()
    elif mode == "pipe":
        wav, wl = pipe.submit(t, tlm, rf, [300] * len(mine), **kw)
        pipe.drain()
    else:
        wav, wl = model.infer_batch(t, tlm, rf, [300] * len(mine), **kw)
    torch.cuda.synchronize()
    print(mode, "rep", rep, "ok", tuple(wav.shape), float(wav.abs().max()), flush=True)
