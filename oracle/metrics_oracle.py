"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of the evaluation epilogue of the reference:

  * per-threshold metrics exactly as anaysis/metrics.py:183-199 computes them — a Python loop over THRESHOLDS with
    `(preds >= t)` and the scikit-learn definitions (precision / recall / F1 with zero_division=0, accuracy,
    matthews_corrcoef).  PINNED: tests/golden/eval_metrics.npz holds the outputs of the unmodified reference function
    `calculate_MORE_metrics` (scikit-learn is installed in the build container), written by oracle/make_golden.py.
  * torchmetrics' binned (thresholds=THRESHOLDS) AUROC / average precision / curves as engine_for_frame_finetuning.py:
    469-488 calls them.  torchmetrics is an un-pinned, un-vendored dependency (INSTALL.md:27) that is NOT installed in
    this image: this part restates the published algorithm of torchmetrics 1.x
    (functional/classification/precision_recall_curve.py `_binary_precision_recall_curve_update/compute`, roc.py
    `_binary_roc_compute`, auroc.py `_binary_auroc_compute`, average_precision.py `_binary_average_precision_compute`)
    and is **parity unpinned**; it is cross-checked against scikit-learn's exact (un-binned) AUROC / AP on the golden data.
"""
import numpy as np

THRESHOLDS = np.arange(0.00, 1.001, 0.01).tolist()  # anaysis/metrics.py:16


def confusion_at(preds, labels, t):
    """(tn, fp, fn, tp) of the prediction `preds >= t` (anaysis/metrics.py:185); preds keeps its dtype (fp32)."""
    b = preds >= t
    y = labels.astype(bool)
    return int((~b & ~y).sum()), int((b & ~y).sum()), int((~b & y).sum()), int((b & y).sum())


def thresholded(preds, labels, thresholds=THRESHOLDS):
    """anaysis/metrics.py:176-199, one threshold at a time."""
    out = {k: [] for k in ("mcc", "precision", "recall", "acc", "f1")}
    counts = []
    for t in thresholds:
        tn, fp, fn, tp = confusion_at(preds, labels, t)
        counts.append((tn, fp, fn, tp))
        n = tn + fp + fn + tp
        out["precision"].append(tp / (tp + fp) if tp + fp else 0.0)
        out["recall"].append(tp / (tp + fn) if tp + fn else 0.0)
        out["acc"].append((tp + tn) / n)
        out["f1"].append(2 * tp / (2 * tp + fp + fn) if 2 * tp + fp + fn else 0.0)
        # sklearn.metrics.matthews_corrcoef on the confusion matrix C = [[tn, fp], [fn, tp]]
        C = np.array([[tn, fp], [fn, tp]], dtype=np.float64)
        t_sum, p_sum = C.sum(axis=1), C.sum(axis=0)
        cov_ytyp = np.trace(C) * n - t_sum @ p_sum
        cov_ypyp = n * n - p_sum @ p_sum
        cov_ytyt = n * n - t_sum @ t_sum
        out["mcc"].append(0.0 if cov_ypyp * cov_ytyt == 0 else float(cov_ytyp / np.sqrt(cov_ytyt * cov_ypyp)))
    return out, np.array(counts, dtype=np.int64)


def _safe_divide(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.where(b != 0, a / np.where(b != 0, b, 1.0), 0.0)


def torchmetrics_binned(preds, labels, thresholds=THRESHOLDS):
    """Binned binary AUROC / AP / ROC / PR of torchmetrics with `thresholds=` given (eff:469-488)."""
    thr32 = np.asarray(thresholds, dtype=np.float64).astype(np.float32)  # torch.tensor(thresholds) next to fp32 preds
    st = np.array([confusion_at(preds, labels, t) for t in thr32], dtype=np.float64)  # [T, (tn, fp, fn, tp)]
    tns, fps, fns, tps = st.T
    tpr = _safe_divide(tps, tps + fns)[::-1]
    fpr = _safe_divide(fps, fps + tns)[::-1]
    auroc = float(np.trapezoid(tpr, fpr)) if hasattr(np, "trapezoid") else float(np.trapz(tpr, fpr))
    precision = np.concatenate([_safe_divide(tps, tps + fps), [1.0]])
    recall = np.concatenate([_safe_divide(tps, tps + fns), [0.0]])
    ap = float(-np.sum((recall[1:] - recall[:-1]) * precision[:-1]))
    return {"auroc": auroc, "ap": ap, "fpr": fpr, "tpr": tpr, "precision": precision, "recall": recall}


def argmax_metrics(probs2, labels):
    """eff:462-468: prediction = torch.max(softmax, 1) (first maximum on ties), then the binary torchmetrics scores."""
    pred = probs2[:, 1] > probs2[:, 0]
    y = labels.astype(bool)
    tn, fp, fn, tp = int((~pred & ~y).sum()), int((pred & ~y).sum()), int((~pred & y).sum()), int((pred & y).sum())
    d = lambda a, b: a / b if b else 0.0  # noqa: E731
    return {"acc": d(tp + tn, len(y)), "recall": d(tp, tp + fn), "precision": d(tp, tp + fp),
            "f1": d(2 * tp, 2 * tp + fp + fn), "confmat": [[tn, fp], [fn, tp]]}


def synthetic_scores(n, seed=0, tie_fraction=0.05):
    """Per-frame logits / labels with the features that break naive implementations: probabilities that fall exactly on
    thresholds (0, 0.5, 1), saturated scores, ties between the two logits, and a class imbalance."""
    rng = np.random.default_rng(seed)
    labels = (rng.random(n) < 0.3).astype(np.int64)
    logits = rng.normal(size=(n, 2)).astype(np.float32) * 2.0
    logits[:, 1] += labels * 1.5
    k = max(1, int(tie_fraction * n))
    logits[:k, 1] = logits[:k, 0]                         # p = 0.5 exactly
    logits[k:2 * k, 1] = logits[k:2 * k, 0] + 40.0        # p = 1.0 exactly in fp32
    logits[2 * k:3 * k, 1] = logits[2 * k:3 * k, 0] - 120.0  # p = 0.0 exactly (underflow)
    return logits, labels
