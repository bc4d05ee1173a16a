"""GPU tier: the two metric drivers, end to end — triad_b200.retrieval.compute_{av,tv}_retrieval_metrics against the
reference's OWN drivers (src/retrieval.py:146-188, :250-292: subset selection, embedding loop, the two N x N python
aggregator loops, recall@k) run on the same stub model / dataset / subset file.  Needs the staged reference
(oracle/_ref, see oracle/build_ref.py); the drop-in imports the reference's `retrieval` module for the caller-side
helpers (select_subset_indices, embed_*_subset), exactly as in a patched checkout."""
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

N_ITEMS, D = 24, 32


class _Model:
    """Stands in for MultiModalModel: deterministic 'encoders' (fixed random projections of the raw inputs)."""

    def __init__(self):
        g = torch.Generator().manual_seed(99)
        self.temperature = torch.nn.Parameter(torch.tensor(1.5, device="cuda"))
        self.wv = torch.randn(12, D, generator=g).cuda()
        self.wa = torch.randn(10, D, generator=g).cuda()
        self.wt = torch.randn(7, D, generator=g).cuda()

    def eval(self):
        return self

    def visual_embedder(self, frames):                       # (B,3,8,8) -> (B,16,D)
        x = frames.reshape(frames.shape[0], 16, 12)
        return x @ self.wv

    def audio_embedder(self, audio):                         # (B,T) zero-padded -> (B,T//10,D)
        B, T = audio.shape
        n = T // 10
        return audio[:, :n * 10].reshape(B, n, 10) @ self.wa

    def text_embedder(self, captions):                       # list of strings -> ((B,Nt,D), mask)
        lens = [len(c.split()) for c in captions]
        Nt = max(lens)
        x = torch.zeros(len(captions), Nt, 7, device="cuda")
        mask = torch.zeros(len(captions), Nt, dtype=torch.int64, device="cuda")
        for b, c in enumerate(captions):
            for t, w in enumerate(c.split()):
                gw = torch.Generator().manual_seed(hash(w) % (2 ** 31))
                x[b, t] = torch.randn(7, generator=gw).cuda()
                mask[b, t] = 1
        return x @ self.wt, mask


class _AVData:
    def __len__(self):
        return N_ITEMS

    def __getitem__(self, idx, apply_augmentation=False):
        g = torch.Generator().manual_seed(1000 + idx)
        base = torch.randn(40, generator=g)
        frames = torch.randn(3, 8, 8, generator=g) * 0.1
        frames.view(-1)[:40] += base                          # the audio and the frames of an item share a pattern
        audio = torch.cat([base, torch.randn(10 * (idx % 3), generator=g) * 0.1])    # ragged lengths
        return {"video_frames": frames, "audio": audio, "video_path": f"item{idx}.mp4"}


class _TVData:
    WORDS = ["red", "dog", "runs", "blue", "car", "tree", "on", "a", "hill", "fast", "cat", "sits"]

    def __len__(self):
        return N_ITEMS

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(2000 + idx)
        n = 2 + idx % 5
        words = [self.WORDS[int(k)] for k in torch.randint(0, len(self.WORDS), (n,), generator=g)]
        return torch.randn(3, 8, 8, generator=g), " ".join(words) + f" w{idx}"


@pytest.fixture()
def ref_retrieval(monkeypatch):
    from oracle import ref_loader
    ref = ref_loader.load()
    if ref is None:
        pytest.skip("oracle/_ref not staged")
    monkeypatch.setitem(sys.modules, "retrieval", ref[1])       # what `import retrieval` finds in a reference checkout
    return ref[1]


def test_av_metric_driver_matches_the_reference(ref_retrieval, tmp_path, monkeypatch):
    from triad_b200 import retrieval as R
    import torch.utils.data as tud
    real_loader = tud.DataLoader
    monkeypatch.setattr(ref_retrieval, "DataLoader", lambda *a, **k: real_loader(*a, **{**k, "num_workers": 0}))
    model, data = _Model(), _AVData()
    subset = str(tmp_path / "subset_av.json")
    want = ref_retrieval.compute_av_retrieval_metrics(model, data, subset, device="cuda")      # writes the subset file
    got = R.compute_av_retrieval_metrics(model, data, subset, device="cuda")                   # reads the same subset
    assert set(got) == set(want) and len(got) == 8
    for k in want:
        assert got[k] == pytest.approx(want[k], abs=1e-12), k
    assert 0.0 < want["A->V_r20"] < 1.0 and want["A->V_r20"] >= want["A->V_r1"]          # neither trivially 0 nor 1


def test_tv_metric_driver_matches_the_reference(ref_retrieval, tmp_path, monkeypatch):
    from triad_b200 import retrieval as R
    import torch.utils.data as tud
    real_loader = tud.DataLoader
    monkeypatch.setattr(ref_retrieval, "DataLoader", lambda *a, **k: real_loader(*a, **{**k, "num_workers": 0}))
    model, data = _Model(), _TVData()
    subset = str(tmp_path / "subset_tv.json")
    want = ref_retrieval.compute_tv_retrieval_metrics(model, data, subset, device="cuda")
    got = R.compute_tv_retrieval_metrics(model, data, subset, device="cuda")
    assert set(got) == set(want) and len(got) == 8
    for k in want:
        assert got[k] == pytest.approx(want[k], abs=1e-12), k
