"""Host side of the compact wire format (wire.py): record conversion, time packing, batching.  No GPU needed: the packed
tensors are rebuilt with the test-only restatement oracle/wire_port.py."""
import numpy as np
import pytest
import torch

from news_recommendation_model_b200 import wire
from news_recommendation_model_b200.synthetic import make_batch
from oracle.wire_port import expand_reference, unpack_time


def records_of(b):
    return [[b.impression_id[i].numpy(), b.user_id[i].numpy(), b.x_history[i].numpy(), b.x_target[i].numpy(), b.x_global[i].numpy(),
             b.label[i].numpy(), b.label_id[i].numpy(), b.empty_num[i].numpy()] for i in range(b.x_history.shape[0])]


def test_pack_time_roundtrip_and_range_checks():
    t = np.array([[0, 0, 0, 0], [3000, 12, 30, 23], [4095, 15, 31, 31], [1, 2, 3, 4]])
    assert np.array_equal(unpack_time(wire.pack_time(t)), t.astype(np.float64))
    for bad in ([4096, 0, 0, 0], [0, 16, 0, 0], [0, 0, 32, 0], [0, 0, 0, 32], [-1, 0, 0, 0], [0.5, 0, 0, 0]):
        with pytest.raises(ValueError):
            wire.pack_time(np.array([bad]))


@pytest.mark.parametrize('exact', [True, False])
def test_from_records_rebuilds_the_packed_tensors(exact):
    b = make_batch(12, 30, 11, seed=8, user_num=50, variable_history=True, variable_candidates=True, fp32_exact=exact)
    # repeat some articles so that de-duplication has something to do
    b.x_history[3, :5, 4:78] = b.x_history[0, :5, 4:78]
    b.x_target[2, 0, 4:78] = b.x_history[0, 0, 4:78]
    ds = wire.from_records(records_of(b))
    d = ds.data
    assert ds.table.rows.dtype == torch.float32 and ds.table.rows.shape[1] == wire.ARTICLE_COLS
    assert not ds.table.rows[0].any() and ds.table.rows[1:].any(dim=1).all()
    assert int(d.hist_article.max()) < ds.table.n and int(d.hist_article.min()) >= 0
    assert d.hist_article[3, 0] == d.hist_article[0, 0]
    xh, xt, xg = expand_reference(ds.table.rows.numpy(), d.hist_article.numpy(), d.hist_time.numpy(), d.hist_click.numpy(),
                                  d.cand_article.numpy(), d.cand_time.numpy())
    # the model reads x.to(float32) (user_invariant_interest_model.py:74-75): equality is required after that cast
    assert torch.equal(xh.float(), b.x_history.float()) and torch.equal(xt.float(), b.x_target.float())
    assert torch.equal(xg.float(), b.x_global.float())
    if exact:
        assert torch.equal(xh, b.x_history) and torch.equal(xt, b.x_target) and torch.equal(xg, b.x_global)
    assert torch.equal(d.label.double(), b.label) and torch.equal(d.empty_num, b.empty_num)
    assert torch.equal(d.user_id, b.user_id) and torch.equal(d.impression_id, b.impression_id)
    # pad rows / pad candidates point at the pad article
    assert ((b.x_history.abs().sum(-1) == 0) == (d.hist_article == 0)).all()
    assert ((b.x_target.abs().sum(-1) == 0) == (d.cand_article == 0)).all()


def test_compact_dataset_batches_cover_every_impression_once():
    t = wire.make_article_table(500, seed=1)
    full = wire.make_compact_batch(t, 23, 7, 5, seed=2, variable_history=True)
    ds = wire.CompactDataset(t, full)
    seen = []
    for cb in ds.batches(8, shuffle=True, seed=3, pin=False):
        assert cb.shape[1:] == (7, 5)
        seen += cb.impression_id.tolist()
    assert sorted(seen) == sorted(full.impression_id.tolist())
    assert sum(1 for _ in ds.batches(8, drop_last=True, pin=False)) == 2
    assert full.input_bytes() == 23 * (7 * 16 + 5 * 12 + 8)


def test_scoring_bucket_plan_covers_every_impression_with_a_pad_column():
    """Host logic of the ragged-aware scoring (scoring.bucket_plan): groups are disjoint, cover the batch, start on even
    positions of the sorted order, and are wide enough for every member's real candidates plus one pad."""
    from news_recommendation_model_b200.scoring import bucket_plan
    g = torch.Generator().manual_seed(5)
    for B, C, groups, min_group in ((1024, 94, 8, 128), (81, 40, 8, 16), (300, 7, 4, 16), (5, 20, 8, 128), (64, 100, 3, 10)):
        n = torch.randint(0, C + 1, (B,), generator=g)
        n[0] = C
        plan = bucket_plan(n, C, groups, min_group)
        seen = torch.cat([idx for idx, _ in plan])
        assert sorted(seen.tolist()) == list(range(B))
        assert len(plan) <= max(1, min(groups, B // min_group))
        pos, prev_max = 0, C
        for idx, width in plan:
            assert pos % 2 == 0 and 1 <= width <= C
            assert int(n[idx].max()) <= prev_max              # sorted by candidate count, descending
            prev_max = int(n[idx].min())
            assert bool(((n[idx] < width) | (n[idx] == C)).all())   # a pad column inside the width unless the row has no pads
            pos += idx.numel()


def test_prefetch_loader_yields_the_same_batches_as_the_synchronous_generator():
    """wire.PrefetchLoader (background thread, fixed ring, ragged last batch) against CompactDataset.batches on the same order."""
    from news_recommendation_model_b200 import wire
    table = wire.make_article_table(300, seed=3)
    full = wire.make_compact_batch(table, 53, 7, 4, seed=5, user_num=20, variable_history=True)
    ds = wire.CompactDataset(table, full)
    for shuffle in (False, True):
        ref = list(ds.batches(8, shuffle=shuffle, seed=11, pin=False))
        loader = wire.PrefetchLoader(ds, 8, shuffle=shuffle, seed=11, depth=3)
        assert len(loader) == len(ref) == 7
        got = []
        for cb in loader:
            got.append(wire.CompactBatch(*[getattr(cb, f).clone() for f in cb.__dataclass_fields__]))   # the slot is recycled after the next request
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            for f in a.__dataclass_fields__:
                assert torch.equal(getattr(a, f), getattr(b, f)), f
        assert got[-1].shape[0] == 53 % 8
    # a second pass over the same loader object works (one pass at a time)
    assert sum(cb.shape[0] for cb in wire.PrefetchLoader(ds, 16, drop_last=True, depth=2)) == 48


def test_processed_data_volumes_round_trip_through_the_zstd_pickle_format(tmp_path):
    """The reference stores a volume as ONE zstd frame of the pickled record list (tool/process_data.py:449-462).  Write two volumes,
    read them back (with `zstandard` if present, else pyarrow's zstd codec), and build the compact dataset over both: same records,
    same dataset as converting the concatenated list directly."""
    import os
    rng = np.random.default_rng(7)
    H, C = 6, 4
    arts = rng.normal(size=(9, 74)).astype(np.float32).astype(np.float64)

    def record(i):
        hist = np.zeros((H, 80)); cand = np.zeros((C, 78)); glob = np.zeros((C, 3))
        nh, nc = int(rng.integers(1, H + 1)), int(rng.integers(1, C + 1))
        for h in range(nh):
            hist[h, 4:78] = arts[rng.integers(0, 9)]; hist[h, 0:4] = [3, 5, 17, 9]; hist[h, 78:80] = rng.random(2)
        for c in range(nc):
            cand[c, 4:78] = arts[rng.integers(0, 9)]; cand[c, 0:4] = [3, 5, 18, 11]; glob[c] = rng.random(3).astype(np.float32)
        label = np.zeros(C); label[0] = 1
        return [np.int64(1000 + i), np.int64(i % 5), hist, cand, glob, label, np.arange(C, dtype=np.float64), np.int64(C - nc)]

    recs = [record(i) for i in range(10)]
    p0, p1 = os.path.join(tmp_path, 'vol0.zst'), os.path.join(tmp_path, 'vol1.zst')
    wire.save_processed_volume(recs[:6], p0)
    wire.save_processed_volume(recs[6:], p1)
    with open(p0, 'rb') as f:
        assert f.read(4) == b'\x28\xb5\x2f\xfd'                      # zstd frame magic: what zstandard.ZstdDecompressor expects
    back = wire.load_processed_volume(p0) + wire.load_processed_volume(p1)
    assert len(back) == 10
    for a, b in zip(recs, back):
        assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b))
    ds0, ds1 = wire.from_records(recs), wire.from_volumes([p0, p1])
    assert torch.equal(ds0.table.rows, ds1.table.rows)
    for f in ds0.data.__dataclass_fields__:
        assert torch.equal(getattr(ds0.data, f), getattr(ds1.data, f)), f
