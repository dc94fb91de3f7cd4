"""BASELINE config 4 inputs ("batch=128 utterances sharded 8 x B200, mixed zh/en tokenizer", L in [30, 70]): 64 pinyin and 64
English sentences tokenised by the REFERENCE's VoiceBpeTokenizer (bpe_tokenizers/voice_tokenizer.py:32-44 over
zh_tokenizer.json / en_tokenizer.json) -> tests/golden/cfg4_mixed.json {items: [{lang, text, ids}]}.  The sentences are
seeded random word sequences (no corpus is available offline).  Run in the build container:  python tests/golden/make_cfg4.py"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from bpe_tokenizers.voice_tokenizer import VoiceBpeTokenizer  # noqa: E402

ZH = ("da4 jia1 hao3 jin1 tian1 lai2 dian3 xiang3 kan4 de5 dong1 xi1 wo3 men5 shi4 zhong1 guo2 ren2 ni3 ta1 zai4 you3 bu4 le5 "
      "yi1 ge4 shang4 xia4 xue2 sheng1 lao3 shi1 peng2 gong1 zuo4 shi2 jian1 xian4 ming2 nian2 yue4 ri4 qi4 hen3 gao1 xing4 "
      "dian4 hua4 che1 zhan4 fei1 ji1 huo3 shui3 shan1 he2 hai3 feng1 yu3 xue3 chun1 qiu1 dong4 bei3 nan2 zhong4 yao4").split()
EN = ("the quick brown fox jumps over a lazy dog hello world this is test of speech synthesis system with many words for every "
      "sentence today tomorrow weather good morning evening please thank you very much we can hear voice clearly now and then "
      "music river mountain city people time year day night water light sound open close small large number").split()
PUNCT_ZH, PUNCT_EN = ["，", "。", "？"], [",", ".", "?"]


def sentence(rng, words, punct, n):
    out = []
    for i in range(n):
        out.append(rng.choice(words))
        if i % 7 == 6 and i + 1 < n:
            out.append(rng.choice(punct[:1]))
    out.append(rng.choice(punct[1:]))
    return " " + " ".join(out) + " "            # api.py:21: text = ' ' + text + ' '


def main():
    rng = random.Random(1234)
    toks = {"zh": VoiceBpeTokenizer("/root/reference/bpe_tokenizers/zh_tokenizer.json"),
            "en": VoiceBpeTokenizer("/root/reference/bpe_tokenizers/en_tokenizer.json")}
    items = []
    for i in range(128):
        lang = "zh" if i % 2 == 0 else "en"
        target = rng.randint(30, 70)
        n = 6
        best = None
        while True:                                # grow the sentence until its token count reaches the drawn target
            txt = sentence(random.Random(1000 * i + n), ZH if lang == "zh" else EN, PUNCT_ZH if lang == "zh" else PUNCT_EN, n)
            ids = toks[lang].encode(txt)
            if len(ids) > 70:
                break
            best = (txt, ids)
            if len(ids) >= target:
                break
            n += 1
        txt, ids = best
        assert 30 <= len(ids) <= 70, (lang, len(ids))
        items.append({"lang": lang, "text": txt, "ids": ids})
    lens = [len(it["ids"]) for it in items]
    print("L range", min(lens), max(lens), "mean", sum(lens) / len(lens), "max id", max(max(it["ids"]) for it in items))
    json.dump({"note": "ids = reference VoiceBpeTokenizer.encode(text); api.py's trailing pad 0 is NOT included", "items": items},
              open(os.path.join(HERE, "cfg4_mixed.json"), "w"), ensure_ascii=False, indent=0)
    print("wrote", os.path.join(HERE, "cfg4_mixed.json"))


if __name__ == "__main__":
    main()
