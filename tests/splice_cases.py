"""Cases of tests/golden/reference_splice.npz (the reference's own AffectGPT.forward :956-1009 run in the build
container by tests/golden/make_reference_splice.py), decoded for the CPU (oracle) and GPU (agb_splice_tokens) suites."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOK = {"frame": 1001, "face": 1002, "audio": 1003, "multi": 1004, "image": 1005}
VOCAB = 1100
STUB_CASES = ["nq1", "nq_many", "other_order", "sample_without_placeholder", "frame_audio_only",
              "err_count", "err_consecutive", "err_missing_face"]
_ORDER = ("frame", "face", "audio", "multi")


def fixture():
    return np.load(os.path.join(GOLD, "reference_splice.npz"))


def embed_ids(ids: torch.Tensor, dim: int, seed: int) -> torch.Tensor:
    """What the reference does before its loop (:959-966): placeholder ids -> 0, then embed_tokens.  The table is the stub
    LLM's seeded nn.Embedding weight (make_reference_splice.embedding_table)."""
    table = torch.randn(VOCAB, dim, generator=torch.Generator().manual_seed(seed))
    tmp = ids.clone()
    for t in TOK.values():
        tmp[tmp == t] = 0
    return table[tmp]


def stub_case(fix, name):
    ids = torch.from_numpy(fix[f"{name}.ids"])
    nqv = [int(x) for x in fix[f"{name}.nq"]]
    present = [bool(x) for x in fix[f"{name}.present"]]
    nq = dict(zip(_ORDER, nqv))
    toks = {k: torch.from_numpy(fix[f"{name}.tok_{k}"]) for k, p in zip(_ORDER, present) if p}
    if not (present[1] and present[2]):                         # forward :941-955: multi needs faces and audios
        toks.pop("multi", None)
    emb = embed_ids(ids, int(fix[f"{name}.dim"]), int(fix[f"{name}.embed_seed"]))
    err = str(fix[f"{name}.error"])
    return {"ids": ids, "nq": nq, "tokens": toks, "embeds": emb,
            "want": torch.from_numpy(fix[f"{name}.inputs_embeds"]) if not err else None,
            "error": json.loads(err) if err else None}
