"""Token layout: positions are bit-exact vs the literal restatement of affectgpt.py:956-1009."""
import pytest
import torch

from affectgpt_b200 import splice as P
from oracle import sampling as O

TOK = {"frame": 1001, "face": 1002, "audio": 1003, "multi": 1004, "image": 1005}


def _prompt(b, L, layout, nq, gen):
    ids = torch.randint(5, 900, (b, L), generator=gen)
    for i in range(b):
        pos = 3 + i
        for name in layout:
            ids[i, pos:pos + nq[name]] = TOK[name]
            pos += nq[name] + 2 + i
    return ids


@pytest.mark.parametrize("nq", [{"frame": 1, "face": 1, "audio": 1, "multi": 1}, {"frame": 4, "face": 4, "audio": 2, "multi": 3}])
def test_splice_matches_reference_order(nq):
    g = torch.Generator().manual_seed(0)
    b, L, D = 3, 64, 16
    ids = _prompt(b, L, ["multi", "audio", "face", "frame"], nq, g)     # base_dataset.py:882-889 prompt order
    emb = torch.randn(b, L, D, generator=g)
    feats = {k: torch.randn(b, nq[k], D, generator=g) for k in nq}
    pos = P.placeholder_positions(ids, TOK, nq, list(feats))
    for i in range(b):
        ref = O.splice_positions(ids[i].tolist(), TOK, nq, list(feats))
        assert {k: int(v[i]) for k, v in pos.items()} == ref


def test_splice_errors_like_reference():
    nq = {"frame": 2, "audio": 1}
    ids = torch.full((1, 16), 7)
    ids[0, 3] = TOK["frame"]                                # count != n_query
    emb = torch.zeros(1, 16, 4)
    feats = {"frame": torch.zeros(1, 2, 4), "audio": torch.zeros(1, 1, 4)}
    with pytest.raises(ValueError):
        P.placeholder_positions(ids, TOK, nq, list(feats))
    with pytest.raises(ValueError):
        O.splice(emb, ids, TOK, nq, feats)
    ids[0, 9] = TOK["frame"]                                # right count, not consecutive
    with pytest.raises(ValueError):
        P.placeholder_positions(ids, TOK, nq, list(feats))
    with pytest.raises(ValueError):
        O.splice(emb, ids, TOK, nq, feats)
    ids2 = torch.full((1, 16), 7)
    ids2[0, 5] = TOK["audio"]
    with pytest.raises(AssertionError):                     # placeholder present, embeds missing
        P.placeholder_positions(ids2, TOK, nq, ["frame"])


def test_per_sample_semantics_and_no_cpu_fallback():
    """The reference loops per sample (:968-1005): a sample without a placeholder is left alone even when another sample
    of the batch carries it."""
    nq = {"frame": 2}
    ids = torch.full((2, 12), 7)
    ids[1, 4:6] = TOK["frame"]
    pos = P.placeholder_positions(ids, TOK, nq, ["frame"])
    assert pos["frame"].tolist() == [-1, 4]
    assert O.splice_positions(ids[0].tolist(), TOK, nq, ["frame"]) == {}
    assert O.splice_positions(ids[1].tolist(), TOK, nq, ["frame"]) == {"frame": 4}
    with pytest.raises(RuntimeError):                       # the copy itself is a CUDA kernel: no CPU path
        P.splice_tokens(torch.zeros(2, 12, 4), ids, TOK, nq, {"frame": torch.zeros(2, 2, 4)})
