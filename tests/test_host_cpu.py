"""CPU tests of the host-side layers around the hot path: CLI surface, metric, dataset reader + the
reference's padding / batching rules."""
import numpy as np
import torch

REFERENCE_FLAGS = [  # add_argument flags of the reference parsers (main_frame_val_text_missing.py:213-252)
    '--dataset', '--train_dataset', '--valid_dataset', '--test_dataset', '--audio_feature', '--text_feature',
    '--video_feature', '--feat4_feature', '--debug', '--test_sets', '--save_root', '--savewhole', '--feat_type',
    '--feat_scale', '--model', '--layers', '--n_classes', '--num_folder', '--model_type', '--full_mse_loss_w',
    '--missing_mse_loss_w', '--text_feat_loss_w', '--text_query_feat_loss_w', '--features_loss_w', '--rnc_loss_w',
    '--lr', '--l2', '--dropout', '--batch_size', '--num_workers', '--epochs', '--seed', '--gpu', '--local_rank']


def test_cli_keeps_every_reference_flag_and_default():
    from sdumc_b200.cli import build_parser
    for inference in (False, True):
        p = build_parser(inference)
        ours = {a.option_strings[0]: a for a in p._actions if a.option_strings}
        assert set(REFERENCE_FLAGS) <= set(ours)
        a = p.parse_args([])
        assert (a.full_mse_loss_w, a.missing_mse_loss_w, a.text_feat_loss_w, a.text_query_feat_loss_w,
                a.features_loss_w, a.rnc_loss_w) == (0.5, 0.5, 0.1, 0.7, 0.1, 0.8)
        assert (a.lr, a.l2, a.batch_size, a.epochs, a.seed, a.layers) == (1e-4, 1e-5, 32, 100, 100, '256,128')
    # the canonical invocation of shell/main_text_missing_icassp.sh parses
    build_parser(False).parse_args(
        "--dataset=CMU-MOSEI --valid_dataset=CMU-MOSEI_valid --test_dataset=CMU-MOSEI_test "
        "--model=wengnet_mosei_mult_views_text_missing --test_sets=test3 --num_workers=4 --audio_feature=a "
        "--text_feature=t --video_feature=v --feat4_feature=f --batch_size=96 --lr=1e-4 --epochs=25 --gpu=0 "
        "--full_mse_loss_w=0.5 --missing_mse_loss_w=0.5 --text_feat_loss_w=0 --text_query_feat_loss_w=0 "
        "--features_loss_w=0.13 --rnc_loss_w=0.5".split())


def test_lr_schedule_matches_reference_lambda():
    from sdumc_b200.trainer import lr_lambda
    ref = lambda epoch: (epoch + 1) / 5 if epoch < 5 else 0.9 ** ((epoch + 1 - 5) // 10)  # noqa: E731  (main…:319-320)
    assert all(abs(lr_lambda(e) - ref(e)) < 1e-15 for e in range(60))


def test_metric_keys_and_values():
    from metric import eval_mosei_metric
    r = eval_mosei_metric([0.5, -1.0, 2.0, 0.1], [1.0, -2.0, 1.5, 0.0])
    assert set(r) >= {"mae", "f1"} and abs(r["mae"] - 0.525) < 1e-12 and r["acc2"] == 1.0


def test_reader_and_collate_follow_reference_padding(tmp_path):
    from sdumc_b200.dataset import Store4F, read_names_labels, read_one_feature
    rng = np.random.default_rng(0)
    names = [f"utt{i}" for i in range(7)]
    feats = ("wav", "txt", "vis", "f4")
    dims = (16, 24, 8, 24)
    lens = {}
    for f, d in zip(feats, dims):
        (tmp_path / f).mkdir()
        for n in names:
            L = int(rng.integers(1, 9))
            lens[(f, n)] = L
            np.save(tmp_path / f / f"{n}.npy", rng.standard_normal((L, d)).astype(np.float32))
    # a visual feature stored as a directory of per-frame vectors (read_data.py:33-37)
    (tmp_path / "vis" / "dirutt").mkdir()
    for k in range(3):
        np.save(tmp_path / "vis" / "dirutt" / f"{k:03d}.npy", np.full((8,), float(k), dtype=np.float32))
    assert read_one_feature(str(tmp_path / "vis"), "dirutt").shape == (3, 8)
    corpus = {n: {"emo": 0, "val": float(i) / 3 - 1} for i, n in enumerate(names)}
    np.savez(tmp_path / "label.npz", train_corpus=corpus, val_corpus=corpus, test_corpus=corpus)
    nm, vals = read_names_labels(str(tmp_path / "label.npz"), "train", exclude=["utt3"])
    assert nm == [n for n in names if n != "utt3"] and len(vals) == 6
    st = Store4F.from_disk(str(tmp_path), feats, nm, vals)
    assert st.dims == dims
    batches = list(st.batches(4))
    assert [len(b[2]) for b in batches] == [4, 2]
    b0, v0, n0 = batches[0]
    for key, f in zip(("audio", "text", "video", "feat4"), feats):
        Lmax = max(lens[(f, n)] for n in n0)
        assert b0[key].shape[1] == Lmax                                   # padded to the BATCH maximum
        for j, n in enumerate(n0):
            L = lens[(f, n)]
            ref = torch.from_numpy(np.load(tmp_path / f / f"{n}.npy")).bfloat16()
            assert torch.equal(b0[key][j, :L], ref) and float(b0[key][j, L:].abs().sum()) == 0.0   # right zero-pad
    # a trailing batch of one sample is merged (the reference model crashes on B == 1)
    st5 = Store4F.from_disk(str(tmp_path), feats, nm[:5], vals[:5])
    assert [len(b[2]) for b in st5.batches(4)] == [5]
    # whole batches are dealt to ranks; lock-step training drops the remainder
    st6 = Store4F.from_disk(str(tmp_path), feats, nm, vals)
    assert [len(b[2]) for b in st6.batches(2, rank=1, world=2)] == [2]
    assert [len(b[2]) for b in st6.batches(2, rank=0, world=2, lockstep=True)] == [2]


def test_device_store_mirrors_the_host_store_batching():
    """DeviceStore4F (packed rows + offsets; here on the CPU device - the collate kernel itself is a GPU test):
    same batch composition, labels, names and per-batch padded lengths as Store4F, for every rank."""
    from sdumc_b200.dataset import STREAMS, DeviceStore4F, Store4F
    host = Store4F.synthetic(23, dims=(16, 24, 8, 24), frames=(12, 5, 9, 5), seed=3, ragged=True)
    dev = DeviceStore4F(host, "cpu")
    for s in STREAMS:
        lens = [int(x.shape[0]) for x in host.feats[s]]
        assert dev.packed[s].shape == (sum(lens), host.feats[s][0].shape[1])
        assert dev.offsets[s].tolist() == [0] + list(np.cumsum(lens))
        i = 7
        assert torch.equal(dev.packed[s][dev.offsets[s][i]:dev.offsets[s][i + 1]], host.feats[s][i])
    for rank, world, lock in ((0, 1, False), (1, 2, False), (0, 2, True)):
        hb = list(host.batches(4, rank, world, lockstep=lock))
        db = list(dev.batches(4, rank, world, lockstep=lock))
        assert len(hb) == len(db) > 0
        for (batch, vals, names), (idx, vals_d, names_d) in zip(hb, db):
            assert names == names_d and torch.equal(vals, vals_d)
            assert dev.batch_frames(idx) == tuple(batch[s].shape[1] for s in STREAMS)


def test_prefetch_iterator_stages_one_batch_ahead():
    """cli.prefetched: batch i+1 is staged (H2D in flight) before batch i is handed to the step, every batch is
    committed exactly once and in order; a DeviceStore4F goes through load_from_store instead."""
    from sdumc_b200.cli import prefetched
    from sdumc_b200.dataset import DeviceStore4F, Store4F

    class FakeTrainer:
        def __init__(self):
            self.log = []

        def stage_batch(self, a, t, v, f, vals):
            self.log.append(("stage", int(a.shape[0])))

        def commit_staged(self):
            self.log.append(("commit",))

        def load_from_store(self, store, idx, labels=True):
            self.log.append(("gather", tuple(idx)))

    host = Store4F.synthetic(10, dims=(16, 24, 8, 24), frames=(6, 3, 4, 3), seed=1)
    tr = FakeTrainer()
    seen = []
    for batch, vals, names in prefetched(tr, host.batches(4), host):
        seen.append(len(names))
        tr.log.append(("step", len(names)))
    assert seen == [4, 4, 2]
    assert tr.log == [("stage", 4), ("commit",), ("stage", 4), ("step", 4), ("commit",), ("stage", 2), ("step", 4),
                      ("commit",), ("step", 2)]
    tr = FakeTrainer()
    dev = DeviceStore4F(host, "cpu")
    got = [idx for idx, vals, names in prefetched(tr, dev.batches(4), dev)]
    assert got == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    assert tr.log == [("gather", (0, 1, 2, 3)), ("gather", (4, 5, 6, 7)), ("gather", (8, 9))]


# ---- golden vectors of the reference's own collate code (oracle/make_golden_collate.py) ----------------------
def _collate_golden():
    from pathlib import Path
    z = np.load(Path(__file__).parent / "golden" / "collate_small.npz")
    streams = ("audio", "text", "video", "feat4")
    n = z["pads"].shape[1]
    feats = {s: [torch.from_numpy(z[f"in/{s}/{i}"]).bfloat16() for i in range(n)] for s in streams}
    return z, feats, n


def test_host_collate_equals_reference_padding_golden():
    """Store4F.collate against the padded batches produced by the reference's pad_to_maxlen_pre_modality_tensor_4
    (read_data.py:223-248) + torch.stack (feat_data.py:242-247), bit for bit (values are bf16-exact)."""
    from sdumc_b200.dataset import Store4F
    z, feats, n = _collate_golden()
    store = Store4F(feats, [0.0] * n, [f"u{i}" for i in range(n)])
    batch, _, _ = store.collate(list(range(n)))
    for s in ("audio", "text", "video", "feat4"):
        assert np.array_equal(batch[s].float().numpy(), z[f"batch/{s}"]), s
        lens = np.array([x.shape[0] for x in feats[s]])
        assert np.array_equal(batch[s].shape[1] - lens, z["pads"][("audio", "text", "video", "feat4").index(s)])


def test_batch_chunks_equal_shards_and_reference_order():
    from sdumc_b200.dataset import batch_chunks
    # single process: the reference's consecutive batches; a trailing single sample joins the last batch
    assert [len(c) for c in batch_chunks(65, 32)] == [32, 33]
    assert batch_chunks(64, 32)[1][0] == 32
    # scoring over ranks: whole reference batches round-robin
    assert [c[0] for c in batch_chunks(200, 32, rank=1, world=2)] == [32, 96, 160]
    # lock-step training: every rank has the same number of steps AND the same shard size at every step
    for n, bs, w in ((1000, 32, 2), (1000, 32, 8), (70, 32, 2), (20480, 512, 8), (33, 32, 4)):
        per = [batch_chunks(n, bs, r, w, lockstep=True) for r in range(w)]
        assert len({len(p) for p in per}) == 1
        for step in range(len(per[0])):
            assert len({len(p[step]) for p in per}) == 1
            assert all(len(p[step]) >= 2 for p in per)
        used = sorted(i for p in per for c in p for i in c)
        assert used == list(range(len(used))) and n - len(used) < 2 * w     # a prefix; fewer than 2*world dropped


def test_kfold_indices_equal_sklearn_kfold():
    from sklearn.model_selection import KFold
    from sdumc_b200.dataset import kfold_indices
    for n, k, seed in ((103, 5, 100), (20480, 5, 100), (17, 3, 7)):
        ref = list(KFold(k, shuffle=True, random_state=seed).split(np.arange(n)))
        for (tr, va), (tr2, va2) in zip(ref, kfold_indices(n, k, seed)):
            assert tr.tolist() == tr2 and va.tolist() == va2


def test_store_subset_shares_tensors():
    from sdumc_b200.dataset import Store4F
    z, feats, n = _collate_golden()
    store = Store4F(feats, list(range(n)), [f"u{i}" for i in range(n)])
    sub = store.subset([3, 1])
    assert len(sub) == 2 and sub.names == ["u3", "u1"] and sub.vals.tolist() == [3.0, 1.0]
    assert sub.feats["audio"][0].data_ptr() == feats["audio"][3].data_ptr()
