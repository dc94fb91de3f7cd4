"""Dump the reference model's state-dict layout (keys, shapes, dtypes, storage aliases).

Run in the build container only (needs /root/reference).  Output: detail_tts_b200/manifest.json,
the checkpoint-format contract of prepare/load_infer.py:8-34 (strict load) that the synthetic
checkpoint generator (detail_tts_b200/synth.py) and the weight pre-packer follow.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(__file__))
import refshim  # noqa: E402


def main():
    model, cfg = refshim.build_reference_model()
    sd = model.state_dict()
    seen = {}
    entries = []
    for k, v in sd.items():
        key = (v.data_ptr(), tuple(v.shape), str(v.dtype))
        alias = seen.get(key) if v.numel() > 0 else None
        if alias is None:
            seen[key] = k
        entries.append({"key": k, "shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", ""),
                        "alias_of": alias})
    out = os.path.join(os.path.dirname(__file__), "..", "..", "detail_tts_b200", "manifest.json")
    with open(out, "w") as f:
        json.dump({"config": cfg, "entries": entries}, f, indent=0)
    print(len(entries), "entries;", sum(1 for e in entries if e["alias_of"]), "aliases")


if __name__ == "__main__":
    main()
