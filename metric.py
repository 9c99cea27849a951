"""`from metric import *` of the reference CLIs (main_frame_val_text_missing.py:39).  The reference calls
eval_mosei_metric (:366-367) but defines it nowhere; this is the standard CMU-MOSEI regression metric set
(MAE, Pearson corr, Acc-2 / F1 on non-zero labels as in toolkit/dataloader/cmumosei.py:149-163, Acc-7)."""
import numpy as np

__all__ = ["eval_mosei_metric"]


def eval_mosei_metric(preds, labels, names=None):
    p = np.asarray(preds, dtype=np.float64).reshape(-1)
    y = np.asarray(labels, dtype=np.float64).reshape(-1)
    mae = float(np.mean(np.abs(p - y)))
    corr = float(np.corrcoef(p, y)[0, 1]) if p.std() > 0 and y.std() > 0 else 0.0
    nz = y != 0
    bp, by = p[nz] > 0, y[nz] > 0
    acc2 = float(np.mean(bp == by)) if nz.any() else 0.0
    f1s, ws = [], []
    for cls in (False, True):                      # weighted F1 over the two classes
        tp = np.sum((bp == cls) & (by == cls))
        fp = np.sum((bp == cls) & (by != cls))
        fn = np.sum((bp != cls) & (by == cls))
        f1s.append(2 * tp / max(2 * tp + fp + fn, 1))
        ws.append(np.sum(by == cls))
    f1 = float(np.dot(f1s, ws) / max(sum(ws), 1))
    acc7 = float(np.mean(np.round(np.clip(p, -3, 3)) == np.round(np.clip(y, -3, 3))))
    return {"mae": mae, "corr": corr, "acc2": acc2, "f1": f1, "acc7": acc7}
