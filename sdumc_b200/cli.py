"""Shared driver of the two drop-in CLIs (main_frame_val_text_missing.py and ..._inference.py).

Every flag of the reference parsers (main_frame_val_text_missing.py:213-252, ..._inference.py:251-288) is
kept with its default; additive flags select the data source because the reference hard-codes dataset
paths in config.py (:8-66): --feat_root / --label_path for the reference's on-disk layout, or --synthetic N
for S0-shaped synthetic utterances.  Multi-GPU: launch with torch.distributed.run (one process per GPU).
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch


def build_parser(inference: bool) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    # ---- reference flags (kept verbatim) ----
    p.add_argument('--dataset', type=str, default=None, help='dataset')
    p.add_argument('--train_dataset', type=str, default=None, help='dataset')
    p.add_argument('--valid_dataset', type=str, default=None, help='dataset')
    p.add_argument('--test_dataset', type=str, default=None, help='dataset')
    p.add_argument('--audio_feature', type=str, default=None, help='audio feature name')
    p.add_argument('--text_feature', type=str, default=None, help='text feature name')
    p.add_argument('--video_feature', type=str, default=None, help='video feature name')
    p.add_argument('--feat4_feature', type=str, default=None, help='4th feature name')
    p.add_argument('--debug', action='store_true', default=False, help='whether use debug to limit samples')
    p.add_argument('--test_sets', type=str, default='test1,test2', help='process on which test sets')
    p.add_argument('--save_root', type=str, default='./saved', help='save prediction results and models')
    p.add_argument('--savewhole', action='store_true', default=False, help='whether save latent embeddings')
    p.add_argument('--feat_type', type=str, default='frm_unalign', help='feature type [utt, frm_align, frm_unalign]')
    p.add_argument('--feat_scale', type=int, default=1, help='pre-compress input')
    p.add_argument('--model', type=str, default='wengnet', help='model name for training')
    p.add_argument('--layers', type=str, default='256,128', help='hidden size in model training')
    p.add_argument('--n_classes', type=int, default=-1)
    p.add_argument('--num_folder', type=int, default=-1)
    p.add_argument('--model_type', type=str, default='mlp')
    p.add_argument('--full_mse_loss_w', type=float, default=0.5)
    p.add_argument('--missing_mse_loss_w', type=float, default=0.5)
    p.add_argument('--text_feat_loss_w', type=float, default=0.1)
    p.add_argument('--text_query_feat_loss_w', type=float, default=0.7)
    p.add_argument('--features_loss_w', type=float, default=0.1)
    p.add_argument('--rnc_loss_w', type=float, default=0.8)
    p.add_argument('--lr', type=float, default=0.0001, metavar='LR')
    p.add_argument('--l2', type=float, default=0.00001, metavar='L2')
    p.add_argument('--dropout', type=float, default=0.5, metavar='dropout')
    p.add_argument('--batch_size', type=int, default=32, metavar='BS')
    p.add_argument('--num_workers', type=int, default=0, metavar='nw')
    p.add_argument('--epochs', type=int, default=100, metavar='E')
    p.add_argument('--seed', type=int, default=100)
    p.add_argument('--gpu', default=0, type=int)
    p.add_argument('--local_rank', default=0, type=int)
    # ---- additive flags ----
    p.add_argument('--feat_root', type=str, default=None, help='root of <feature_name>/<utt>.npy (config.PATH_TO_FEATURES)')
    p.add_argument('--label_path', type=str, default=None, help='label .npz (config.PATH_TO_LABEL)')
    p.add_argument('--exclude_names', type=str, default=None,
                   help='text file of train utterances to drop (the reference drops 51 over-long ones, cmumosei.py:10-62)')
    p.add_argument('--synthetic', type=int, default=0, help='N synthetic S0-shaped utterances per split instead of files')
    p.add_argument('--synthetic_ragged', action='store_true', help='ragged synthetic lengths (padding semantics)')
    p.add_argument('--device_store', action='store_true',
                   help='keep the feature store in HBM and build batches with the collate kernel (no per-step H2D)')
    p.add_argument('--folds', type=int, default=1,
                   help='K > 1: K-fold cross-validation over the train split (KFold, shuffle, random_state=--seed; a fresh '
                        'model / optimizer / schedule per fold like the reference loop :295-321); 1 = the reference\'s '
                        'single train/val/test run')
    p.add_argument('--synthetic_frames', type=str, default=None, help='La,Lt,Lv,L4 of --synthetic (default S0: 384,64,256,64)')
    p.add_argument('--synthetic_dims', type=str, default=None, help='Da,Dt,Dv,D4 of --synthetic (default S0: 1024,4096,1024,4096)')
    p.add_argument('--synthetic_label_scale', type=float, default=1.0,
                   help='multiplies the labels of --synthetic (the reference keeps a checkpoint only below a test MAE of 1.0, :299)')
    p.add_argument('--general_dim', type=int, default=256,
                   help='model width (the reference hard-codes general_dim = 256, :191); 1024 = the "hidden 1024" stress config')
    p.add_argument('--no_dropout', action='store_true', help='train without dropout (deterministic parity runs)')
    p.add_argument('--checkpoint', type=str, default=None,
                   help='checkpoint with ["state_dict"] (inference; the reference hard-codes its path, ..._inference.py:341)')
    p.add_argument('--save_checkpoints', action='store_true', help='torch.save the best epochs (commented out in the reference :375)')
    return p


def _dist():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = int(os.environ.get("RANK", "0"))
    return world, rank, (dist.group.WORLD if world > 1 else None)


def load_splits(args):
    from .dataset import Store4F, read_names_labels
    if args.synthetic:
        dims = tuple(int(x) for x in args.synthetic_dims.split(',')) if args.synthetic_dims else (1024, 4096, 1024, 4096)
        frames = tuple(int(x) for x in args.synthetic_frames.split(',')) if args.synthetic_frames else (384, 64, 256, 64)
        def mk(n, seed):
            st_ = Store4F.synthetic(n, dims, frames, seed=seed, ragged=args.synthetic_ragged)
            st_.vals = st_.vals * args.synthetic_label_scale
            return st_
        return mk(args.synthetic, 1234), mk(max(args.batch_size, args.synthetic // 8), 4321), \
            mk(max(args.batch_size, args.synthetic // 8), 9876)
    if not (args.feat_root and args.label_path):
        raise SystemExit("give --feat_root and --label_path (reference on-disk layout) or --synthetic N")
    names_feats = (args.audio_feature, args.text_feature, args.video_feature, args.feat4_feature)
    ex = [l.strip() for l in open(args.exclude_names)] if args.exclude_names else []
    stores = []
    for split in ("train", "val", "test"):
        names, vals = read_names_labels(args.label_path, split, args.debug, ex if split == "train" else ())
        print(f'{split}: sample number {len(names)}')
        stores.append(Store4F.from_disk(args.feat_root, names_feats, names, vals))
    return tuple(stores)


def prefetched(tr, batches, store=None, labels: bool = True):
    """Iterates (batch, vals, names) with the NEXT batch's host->device copy already in flight on the trainer's
    copy stream (the reference gets the same overlap from DataLoader workers + pin_memory).  A DeviceStore4F
    yields index lists instead: the batch is gathered on the device by the collate kernel."""
    from .dataset import DeviceStore4F
    if isinstance(store, DeviceStore4F):
        for idx, vals, nm in batches:
            tr.load_from_store(store, idx, labels=labels)
            yield idx, vals, nm
        return
    it = iter(batches)
    nxt = next(it, None)
    if nxt is not None:
        tr.stage_batch(nxt[0]["audio"], nxt[0]["text"], nxt[0]["video"], nxt[0]["feat4"], nxt[1])
    while nxt is not None:
        cur = nxt
        tr.commit_staged()
        nxt = next(it, None)
        if nxt is not None:
            tr.stage_batch(nxt[0]["audio"], nxt[0]["text"], nxt[0]["video"], nxt[0]["feat4"], nxt[1])
        yield cur


def run_split(tr, store, batch_size, train: bool, rank, world):
    """train_or_eval_model (main…:74-178): one pass over a split.  Returns the reference's result dict."""
    preds_full, preds_missing, labels, names = [], [], [], []
    for batch, vals, nm in prefetched(tr, store.batches(batch_size, rank, world, lockstep=train), store):
        if train:
            tr.train_step()
            pf, pm = tr.predictions()
        else:
            out = tr.score()
            pf, pm = out["val_preds_full"], out["val_preds_missing"]
        preds_full.append(pf.float().cpu().numpy())           # the reference syncs per step too (:156-158)
        preds_missing.append(pm.float().cpu().numpy())
        labels.append(vals.numpy())
        names += nm
    pf, pm, y = np.concatenate(preds_full), np.concatenate(preds_missing), np.concatenate(labels)
    return {"val_mse_full": float(np.mean((pf.reshape(-1) - y) ** 2)),
            "val_mse_missing": float(np.mean((pm.reshape(-1) - y) ** 2)),
            "val_preds_full": pf, "val_preds_missing": pm, "val_labels": y, "names": names}


def make_trainer(args, stores, device, pg, state_dict=None):
    """One trainer serves every split: its static buffers are sized for the longest utterances of all of them."""
    from .trainer import Trainer
    store = stores[0]
    max_frames = tuple(max(s.max_frames[i] for s in stores) for i in range(4))
    w = (args.full_mse_loss_w, args.missing_mse_loss_w, args.text_feat_loss_w, args.text_query_feat_loss_w,
         args.features_loss_w, args.rnc_loss_w)
    # +1 only when some split really ends with a single sample that joins its last batch (batch_chunks); the step is
    # captured per recurring batch shape (Trainer.train_step), so regular batches replay a graph either way
    merge = any(len(s) > args.batch_size and len(s) % args.batch_size == 1 for s in stores)
    cap = max(args.batch_size + (1 if merge else 0), 2)
    tr = Trainer(store.dims, cap, max_frames, device, lr=args.lr, weight_decay=args.l2, loss_w=w,
                 seed=args.seed, process_group=pg, state_dict=state_dict, use_graph=True,
                 general_dim=getattr(args, "general_dim", 256))
    tr.train_dropout = not getattr(args, "no_dropout", False)
    return tr


def main_train(argv=None):
    from metric import eval_mosei_metric
    from .trainer import lr_lambda
    args = build_parser(False).parse_args(argv)
    args.n_classes, args.num_folder = 6, 5                    # reference :255-256
    args.test_sets = args.test_sets.split(',')
    world, rank, pg = _dist()
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        print(args)
        print('====== Reading Data =======')
    train, val, test = load_splits(args)
    if args.device_store:
        from .dataset import DeviceStore4F
        train, val, test = (DeviceStore4F(st_, device) for st_ in (train, val, test))
    args.input_dims = train.dims
    if rank == 0:
        print('====== Training and Evaluation =======')
    best_valid = {'mae': 1.0, 'f1': 0}
    best_full, best_missing = {'mae': 1.0, 'f1': 0}, {'mae': 1.0, 'f1': 0}
    if args.folds > 1:
        from .dataset import kfold_indices
        folds = [(train.subset(tr_i), train.subset(va_i)) for tr_i, va_i in kfold_indices(len(train), args.folds, args.seed)]
    else:
        folds = [(train, val)]
    fold_val = []                                             # per fold: val MSE (full, missing) after the last epoch
    saved = []
    n_steps = n_replays = 0
    for ii, (train_f, val_f) in enumerate(folds):
        if rank == 0:
            print(f'>>>>> Cross-validation: training on the {ii+1} folder >>>>>')
            print('Step1: build model (each folder has its own model)')
        start = time.time()
        torch.manual_seed(args.seed + ii)
        tr = make_trainer(args, (train_f, val_f, test), device, pg)
        if rank == 0:
            print('Step2: training (multiple epoches)')
        vres = None
        for epoch in range(args.epochs):
            tr.set_lr(args.lr * lr_lambda(epoch))             # LambdaLR, stepped once per epoch (:318-321,:342)
            t0 = time.time()
            tres = run_split(tr, train_f, args.batch_size, True, rank, world)
            if rank == 0:
                print('used: {} s'.format(time.time() - t0))
            vres = run_split(tr, val_f, args.batch_size, False, 0, 1)
            if rank == 0:
                print('epoch:%d; train_val_mse_full:%.4f; train_val_mse_missing:%.4f' %
                      (epoch + 1, tres['val_mse_full'], tres['val_mse_missing']))
                if args.folds > 1:
                    print('epoch:%d; fold:%d; val_mse_full:%.4f; val_mse_missing:%.4f' %
                          (epoch + 1, ii + 1, vres['val_mse_full'], vres['val_mse_missing']))
            te = run_split(tr, test, args.batch_size, False, 0, 1)
            r_full = eval_mosei_metric(te['val_preds_full'], te['val_labels'], te['names'])
            r_miss = eval_mosei_metric(te['val_preds_missing'], te['val_labels'], te['names'])
            if r_full['mae'] <= best_full['mae']:
                best_full = dict(r_full, epoch=epoch)
                if rank == 0:
                    print("***************better full**********************")
                    if args.save_checkpoints:                 # the reference's commented-out torch.save (:375)
                        os.makedirs(args.save_root, exist_ok=True)
                        path = os.path.join(args.save_root, f'mosei_mult-view_kd_full_{best_full["mae"]}_{epoch+1}.pt')
                        torch.save(tr.checkpoint(epoch + 1), path)
                        saved.append(path)
            if r_miss['mae'] <= best_missing['mae']:
                best_missing = dict(r_miss, epoch=epoch)
                if rank == 0:
                    print("===============better missing===================")
                    if args.save_checkpoints:                 # (:384)
                        os.makedirs(args.save_root, exist_ok=True)
                        path = os.path.join(args.save_root, f'mosei_mult-view_kd_missing_{best_missing["mae"]}_{epoch+1}.pt')
                        torch.save(tr.checkpoint(epoch + 1), path)
                        saved.append(path)
            if rank == 0:
                print("test full:")
                print(r_full)
                print("test missing:")
                print(r_miss)
                print("-" * 50)
        fold_val.append((vres['val_mse_full'], vres['val_mse_missing']) if vres else None)
        n_steps, n_replays = n_steps + tr.n_steps, n_replays + tr.n_replays
        tr.close()
        if rank == 0:
            print(f'>>>>> Finish: training on the {ii+1} data, duration: {time.time() - start} >>>>>')
    if rank == 0:
        print('====== Gain predition on test data =======')
        print("best_valid:")
        print(best_valid)
        print("best_test_full:")
        print(best_full)
        print("best_test_missing:")
        print(best_missing)
        with open('features_ablation_study.txt', mode='a') as f:   # reference :411-416
            f.write(f'--full_mse_loss_w={args.full_mse_loss_w} --missing_mse_loss_w={args.missing_mse_loss_w} '
                    f'--text_feat_loss_w={args.text_feat_loss_w} --text_query_feat_loss_w={args.text_query_feat_loss_w} '
                    f'--features_loss_w={args.features_loss_w} --rnc_loss_w={args.rnc_loss_w}\n')
            f.write(str(best_full) + '\n' + str(best_missing) + '\n')
        print(f'{args.audio_feature}+{args.text_feature}+{args.video_feature}')
    return {"fold_val_mse": fold_val, "best_test_full": best_full, "best_test_missing": best_missing,
            "checkpoints": saved, "train_steps": n_steps, "graph_replays": n_replays}


def gather_results(res: dict, order: np.ndarray, n_total: int, pg, device):
    """All ranks' per-utterance result arrays -> the full arrays in dataset order on every rank.
    `order` = dataset positions of this rank's rows (ranks hold whole reference batches, dealt round-robin).
    One padded all_gather per key over NCCL (the payload of 100k utterances is ~1.1 GB in total)."""
    import torch.distributed as dist
    world = dist.get_world_size(pg)
    n_loc = torch.tensor([len(order)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(counts, n_loc, group=pg)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)

    def gather(arr: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
        pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
        pad[: t.shape[0]] = t
        out = torch.empty((world,) + tuple(pad.shape), dtype=t.dtype, device=device)
        dist.all_gather_into_tensor(out, pad, group=pg)
        return out.cpu().numpy()
    pos = gather(order.astype(np.int64))
    full = {}
    for k, v in res.items():
        if not isinstance(v, np.ndarray):
            continue
        g = gather(v)
        dst = np.zeros((n_total,) + v.shape[1:], dtype=v.dtype)
        for r in range(world):
            dst[pos[r, : counts[r]]] = g[r, : counts[r]]
        full[k] = dst
    return full


def main_inference(argv=None):
    args = build_parser(True).parse_args(argv)
    args.test_sets = args.test_sets.split(',')
    world, rank, pg = _dist()
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    train, val, test = load_splits(args)
    if args.device_store:
        from .dataset import DeviceStore4F
        train, val, test = (DeviceStore4F(st_, device) for st_ in (train, val, test))
    sd = None
    if args.checkpoint:
        ck = torch.load(args.checkpoint, map_location="cpu")
        sd = {k.replace('module.', ''): v for k, v in ck['state_dict'].items()}   # ..._inference.py:341 (strict=False)
    torch.manual_seed(args.seed)
    tr = make_trainer(args, (train, val, test), device, None, state_dict=sd)
    keys = ("val_preds_full", "val_preds_missing", "full_rep", "missing_rep", "full_rnc", "missing_rnc",
            "text_rep_query_full", "text_rep_query_missing", "text_rep_full", "text_rep_missing")
    results = {}
    for split_name, store in (("train", train), ("val", val), ("test", test)):
        acc = []                                    # one packed [b, width] row block per batch (one D2H copy each)
        labels, names = [], []
        t0 = time.time()
        for batch, vals, nm in prefetched(tr, store.batches(args.batch_size, rank, world), store, labels=False):  # whole reference batches per rank
            tr.score()
            acc.append(tr.last_packed.cpu())
            labels.append(vals.numpy())
            names += nm
        packed = torch.cat(acc).numpy() if acc else np.zeros((0, 1), dtype=np.float32)
        res = {k: np.ascontiguousarray(v) for k, v in tr.unpack_scores(packed).items()} if acc else {k: packed[:0] for k in keys}
        res["val_labels"], res["names"] = np.concatenate(labels), names
        if world > 1:
            # every rank scored whole reference batches (a sample's output depends on its batch's padding); put the
            # rows back in dataset order on every rank, as main_frame_val_text_missing_inference.py:199-214 holds them
            from .dataset import batch_chunks
            order = np.array([i for c in batch_chunks(len(store), args.batch_size, rank, world) for i in c], dtype=np.int64)
            all_names = list(store.names) if getattr(store, "_ids", None) is None else [store.names[i] for i in store._ids]
            res = gather_results(res, order, len(store), pg, device)
            res["names"] = all_names
        res["val_mse_full"] = float(np.mean((res["val_preds_full"].reshape(-1) - res["val_labels"]) ** 2))
        res["val_mse_missing"] = float(np.mean((res["val_preds_missing"].reshape(-1) - res["val_labels"]) ** 2))
        results[split_name] = res
        print(f'[rank {rank}] {split_name}: {len(names)} utterances in {time.time() - t0:.2f} s; '
              f'val_mse_full:{res["val_mse_full"]:.4f}; val_mse_missing:{res["val_mse_missing"]:.4f}')
    if args.savewhole:
        os.makedirs(args.save_root, exist_ok=True)
        np.savez(os.path.join(args.save_root, f'inference_rank{rank}.npz'),
                 **{f"{s}/{k}": v for s, r in results.items() for k, v in r.items() if k != "names"})
    return results
