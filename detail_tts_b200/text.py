"""Text front-end of api.py (second half): BPE token ids for a BATCH of utterances.

Mirrors the reference's call surface:
  VoiceBpeTokenizer(vocab_file).encode / decode / preprocess_text      bpe_tokenizers/voice_tokenizer.py:32-57
  api.py:19-26: pinyin string -> ids -> IntTensor -> F.pad(ids, (0, 1))
and adds what a batched `infer_batch` needs: `encode_batch` (ragged ids padded to one [B, Lmax+1] int32 tensor + lengths that
count api.py's trailing pad) and a language router for mixed zh / en batches (BASELINE config 4: both vocabularies share the
model's single 256-entry text embedding, gpt/model.py:308).  The arithmetic (BPE merges) lives in the third-party
`tokenizers` package, exactly as in the reference; the vocabulary files are the reference's `bpe_tokenizers/*.json` (data, not
shipped here: pass their path).  Chinese text must already be TONE3 pinyin (`lazy_pinyin(..., Style.TONE3,
neutral_tone_with_five=True)`, api.py:20): pypinyin is not a dependency of this package.
"""
import re

import torch

_PUNCT = {"{": "(", "}": ")", "[": "(", "]": ")", "`": "'", "—": "-", "ʼ": "'"}
_PUNCT_RE = re.compile("|".join(re.escape(k) for k in sorted(_PUNCT, key=len, reverse=True)))
_EXTRANEOUS_RE = re.compile(r"^[@#%_=\$\^&\*\+\\]$")


def remove_extraneous_punctuation(word):
    """bpe_tokenizers/voice_tokenizer.py:14-29: bracket / quote / dash normalisation; a lone symbol word is dropped."""
    word = _PUNCT_RE.sub(lambda m: _PUNCT[m.group(0)], word)
    return _EXTRANEOUS_RE.sub("", word)


class VoiceBpeTokenizer:
    def __init__(self, vocab_file):
        from tokenizers import Tokenizer
        self.tokenizer = Tokenizer.from_file(vocab_file) if vocab_file is not None else None

    def preprocess_text(self, txt):
        return remove_extraneous_punctuation(txt)

    def encode(self, txt):
        """voice_tokenizer.py:41-44"""
        txt = self.preprocess_text(txt).replace(" ", "[SPACE]")
        return self.tokenizer.encode(txt).ids

    def decode(self, seq):
        """voice_tokenizer.py:46-54"""
        if isinstance(seq, torch.Tensor):
            seq = seq.cpu().numpy()
        txt = self.tokenizer.decode(seq, skip_special_tokens=False).replace(" ", "")
        return txt.replace("[SPACE]", " ").replace("[STOP]", "").replace("[UNK]", "")

    def encode_batch(self, texts, wrap_spaces=True):
        """api.py:21-25 for every utterance of a batch, through the tokenizers library's batch entry point.
        Returns (ids [B, Lmax + 1] int32 zero-padded, lengths [B] counting api.py's trailing pad id 0)."""
        pre = [self.preprocess_text((" " + t.strip() + " ") if wrap_spaces else t).replace(" ", "[SPACE]") for t in texts]
        return pad_ids([e.ids for e in self.tokenizer.encode_batch(pre)])


def pad_ids(id_lists):
    """Ragged id lists -> ([B, Lmax + 1] int32, lengths [B]): every row carries api.py:25's `F.pad(text_tokens, (0, 1))` pad,
    which the model counts as part of the text (SynthesizerTrn.infer receives text_lengths = L + 1)."""
    lens = [len(x) + 1 for x in id_lists]
    out = torch.zeros(len(id_lists), max(lens), dtype=torch.int32)
    for b, ids in enumerate(id_lists):
        out[b, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return out, lens


class MixedTokenizer:
    """One tokenizer per language tag over the model's single text-embedding id space: `encode_batch(texts, langs)`."""

    def __init__(self, vocab_files):
        self.tok = {lang: VoiceBpeTokenizer(path) for lang, path in vocab_files.items()}

    def encode_batch(self, texts, langs, wrap_spaces=True):
        ids = [None] * len(texts)
        for lang in sorted(set(langs)):
            sel = [i for i, l in enumerate(langs) if l == lang]
            t = self.tok[lang]
            pre = [t.preprocess_text((" " + texts[i].strip() + " ") if wrap_spaces else texts[i]).replace(" ", "[SPACE]") for i in sel]
            for i, e in zip(sel, t.tokenizer.encode_batch(pre)):
                ids[i] = e.ids
        return pad_ids(ids)
