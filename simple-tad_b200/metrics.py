"""Evaluation epilogue of frame-level inference — the metrics `final_test` computes after the score gather
(engine_for_frame_finetuning.py:448-497), the per-threshold tables of anaysis/metrics.py:133-207 and the exact
(un-binned) group summaries of anaysis/metrics.py:19-125.

The reduction over the n per-frame scores runs on the device (`stad_eval_hist`: one pass, exact integer counts against
the 101 THRESHOLDS); what is left is arithmetic on a [2, 102] integer table, done here in float64.  The counts are
additive, so under torchrun the ranks all-reduce that table (808 + 32 bytes) instead of gathering every prediction the
way `gather_predictions_nontensor` (utils.py:791-810) does; the gathered logits are only needed for predictions.csv.

Definitions follow the libraries the reference calls:
  * torchmetrics (binary, `thresholds=THRESHOLDS`, eff:469-488): confusion matrices at `p >= t`; ROC = (fpr, tpr)
    flipped to ascending fpr; AUROC = trapezoid; PR curve with the (precision 1, recall 0) end point; AP =
    -sum((r[1:] - r[:-1]) * p[:-1]).  torchmetrics is not pinned by the reference (INSTALL.md:27) and not installed
    here: this is a restatement of its published binned algorithm (v1.x), see DESIGN.md.
  * sklearn (anaysis/metrics.py:183-199): precision / recall / F1 with zero_division=0, accuracy, Matthews
    correlation at every threshold.
"""
import csv

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

THRESHOLDS = np.arange(0.00, 1.001, 0.01).tolist()  # anaysis/metrics.py:16 (eff:22)


def threshold_tensor(device, thresholds=THRESHOLDS):
    """The thresholds as the fp32 values the reference's `probs >= t` comparison sees for fp32 probabilities."""
    return torch.tensor(np.asarray(thresholds, dtype=np.float64).astype(np.float32), device=device)


def counts_from_hist(hist):
    """hist int64 [2, T+1] (label, number of thresholds <= p) -> dict of int64 [T] arrays tn, fp, fn, tp for the
    predictions `p >= t_k`: positive at threshold k  <=>  bin > k."""
    h = np.asarray(hist, dtype=np.int64)
    above = h[:, ::-1].cumsum(axis=1)[:, ::-1]          # above[y, b] = sum_{b' >= b} h[y, b']
    fp, tp = above[0, 1:], above[1, 1:]                 # bin > k  <=>  bin >= k + 1
    n_neg, n_pos = h[0].sum(), h[1].sum()
    return {"tn": n_neg - fp, "fp": fp, "fn": n_pos - tp, "tp": tp}


def _div(a, b, zero=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.full(np.broadcast(a, b).shape, zero, dtype=np.float64)
    np.divide(a, b, out=out, where=b != 0)
    return out


def thresholded_metrics(c):
    """MCC / precision / recall / accuracy / F1 at every threshold (anaysis/metrics.py:183-199; sklearn definitions with
    zero_division=0; MCC = 0 when a marginal is empty, as sklearn.metrics.matthews_corrcoef returns)."""
    tn, fp, fn, tp = (c[k].astype(np.float64) for k in ("tn", "fp", "fn", "tp"))
    n = tn + fp + fn + tp
    # sklearn's covariance form: cov_ytyp / sqrt(cov_ytyt * cov_ypyp) on the 2x2 confusion matrix
    t_pos, t_neg, p_pos, p_neg = tp + fn, tn + fp, tp + fp, tn + fn
    cov_ytyp = (tp + tn) * n - (t_pos * p_pos + t_neg * p_neg)
    cov_ypyp = n * n - (p_pos * p_pos + p_neg * p_neg)
    cov_ytyt = n * n - (t_pos * t_pos + t_neg * t_neg)
    return {"mcc": _div(cov_ytyp, np.sqrt(cov_ytyt * cov_ypyp)), "precision": _div(tp, tp + fp),
            "recall": _div(tp, tp + fn), "acc": _div(tp + tn, n), "f1": _div(2 * tp, 2 * tp + fp + fn)}


def binned_curves(c, thresholds=THRESHOLDS):
    """torchmetrics' binned ROC / PR curves and their summaries (eff:469-488)."""
    tn, fp, fn, tp = (c[k].astype(np.float64) for k in ("tn", "fp", "fn", "tp"))
    thr = np.asarray(thresholds, dtype=np.float64)
    tpr, fpr = _div(tp, tp + fn)[::-1], _div(fp, fp + tn)[::-1]
    auroc = float(np.sum((fpr[1:] - fpr[:-1]) * (tpr[1:] + tpr[:-1]) / 2.0))
    precision = np.concatenate([_div(tp, tp + fp), [1.0]])
    recall = np.concatenate([_div(tp, tp + fn), [0.0]])
    ap = float(-np.sum((recall[1:] - recall[:-1]) * precision[:-1]))
    return {"auroc": auroc, "ap": ap, "roc_curve": (fpr, tpr, thr[::-1].copy()), "pr_curve": (precision, recall, thr)}


def argmax_metrics(conf):
    """Accuracy / recall / precision / F1 / confusion matrix of the arg-max prediction (eff:464-468)."""
    tn, fp, fn, tp = (float(v) for v in conf)
    n = tn + fp + fn + tp
    d = lambda a, b: a / b if b else 0.0  # noqa: E731
    return {"acc": d(tp + tn, n), "recall": d(tp, tp + fn), "precision": d(tp, tp + fp), "f1": d(2 * tp, 2 * tp + fp + fn),
            "confmat": [[int(tn), int(fp)], [int(fn), int(tp)]]}


def reduce_counts(hist, conf, group=None):
    """Sum the per-rank count tables (int64 [2, T+1] and [4]) over the process group: ONE all-reduce of 8 (2T + 6)
    bytes replaces the gather of every prediction (ut:791-810).  No-op without an initialised group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = torch.cat([hist.flatten(), conf])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        hist, conf = packed[:-4].view_as(hist), packed[-4:]
    return hist, conf


@torch.no_grad()
def evaluate(probs, labels, thresholds=THRESHOLDS, group=None):
    """probs [n, 2] fp32 CUDA (softmax of the logits), labels [n] (any integer dtype) -> dict with the metrics of
    `final_test` (mAP, auroc, acc, P/R/F1 @ arg-max, confmat, binned PR / ROC curves) and the per-threshold lists of
    anaysis/metrics.py.  Under an initialised process group every rank passes ITS shard and receives the metrics of
    the whole set (one all-reduce of the count tables)."""
    if not probs.is_cuda:
        raise RuntimeError("simple-tad_b200 runs on a CUDA (sm_100a) device only; got CPU scores")
    probs = probs.float().contiguous()
    labels = labels.to(device=probs.device, dtype=torch.int32).contiguous()
    if probs.shape[0] > 0:
        hist, conf = _lib.eval_hist(probs, labels, threshold_tensor(probs.device, thresholds))
    else:  # an empty shard (more ranks than windows) still takes part in the reduction below
        hist = torch.zeros(2, len(thresholds) + 1, dtype=torch.int64, device=probs.device)
        conf = torch.zeros(4, dtype=torch.int64, device=probs.device)
    hist, conf = reduce_counts(hist, conf, group)
    c = counts_from_hist(hist.cpu().numpy())
    res = {"counts": c, "n": int(hist.sum())}
    res.update(argmax_metrics(conf.cpu().numpy()))
    res.update(binned_curves(c, thresholds))
    res["thresholded"] = thresholded_metrics(c)
    return res


# ---- exact (un-binned) summaries of anaysis/metrics.py:19-125, on the host ----------------------------------------------
# The analysis scripts of the reference re-read predictions.csv and call scikit-learn on every score (no threshold grid),
# which needs the scores in sorted order; that is a host-side, once-per-evaluation step here as well (numpy, float64).

def _distinct_score_counts(p_risk, labels):
    """Cumulative (true positives, false positives) at the last element of every run of equal scores, scores descending
    (what sklearn's _binary_clf_curve returns), and the class totals."""
    p = np.asarray(p_risk, dtype=np.float64)
    y = np.asarray(labels).astype(np.int64)
    order = np.argsort(-p, kind="mergesort")
    ps, ys = p[order], y[order]
    last = np.r_[np.nonzero(np.diff(ps))[0], y.size - 1]
    tps = np.cumsum(ys)[last].astype(np.float64)
    return tps, (1 + last) - tps, int(y.sum()), int(y.size - y.sum())


def exact_ap(p_risk, labels):
    """average_precision_score (anaysis/metrics.py:54): one PR point per DISTINCT score (ties enter together),
    AP = sum_k (R_k - R_{k-1}) P_k; 0 when there is no positive (sklearn sets recall to one and warns)."""
    tps, fps, n_pos, _ = _distinct_score_counts(p_risk, labels)
    if n_pos == 0:
        return 0.0
    return float(np.sum(np.diff(np.r_[0.0, tps / n_pos]) * (tps / (tps + fps))))


def exact_auroc(p_risk, labels):
    """roc_auc_score (anaysis/metrics.py:56): trapezoid area under the ROC curve with one point per distinct score.
    ValueError when only one class is present."""
    tps, fps, n_pos, n_neg = _distinct_score_counts(p_risk, labels)
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    tpr, fpr = np.r_[0.0, tps / n_pos], np.r_[0.0, fps / n_neg]
    return float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0))


def exact_auroc_ap(p_risk, labels):
    return exact_auroc(p_risk, labels), exact_ap(p_risk, labels)


def calculate_metrics(probs, labels):
    """anaysis/metrics.py:19-64: (acc, precision, recall, f1) of `probs >= 0.5`, exact mAP, exact AUROC.  A group that
    holds one class only gets auc = -10 - label, the reference's fallback (anaysis/metrics.py:57-63); note that the
    scikit-learn installed in this image (1.9) no longer raises there but warns and returns nan, so the unmodified
    reference run here yields nan for such a group — the written fallback is what is implemented."""
    p = np.asarray(probs, dtype=np.float64)
    y = np.asarray(labels).astype(np.int64)
    pred = (p >= 0.5).astype(np.int64)
    tp = float(np.sum((pred == 1) & (y == 1)))
    fp = float(np.sum((pred == 1) & (y == 0)))
    fn = float(np.sum((pred == 0) & (y == 1)))
    tn = float(np.sum((pred == 0) & (y == 0)))
    d = lambda a, b: a / b if b else 0.0  # noqa: E731
    acc, precision, recall, f1 = d(tp + tn, tp + tn + fp + fn), d(tp, tp + fp), d(tp, tp + fn), d(2 * tp, 2 * tp + fp + fn)
    ap = exact_ap(p, y)
    try:
        auc = exact_auroc(p, y)
    except ValueError:
        classes = sorted(set(y.tolist()))
        assert len(classes) == 1
        auc = -10 - classes[0]   # -10: only "normal" frames in the group, -11: only "abnormal"
    return acc, precision, recall, f1, ap, auc


def calculate_fn_group(probs, labels):
    """anaysis/metrics.py:67-92: share of a (all-positive) group scored below 0.5."""
    return float(np.sum(np.asarray(probs, dtype=np.float64) < 0.5) / len(labels))


def calculate_fn_group_thresholds(probs, labels, thresholds=THRESHOLDS):
    """anaysis/metrics.py:95-125: that share at every threshold.  Equals fn / (fn + tp) of `evaluate(...)["counts"]` for
    an all-positive group."""
    p = np.asarray(probs)
    return [float(np.sum(~(p >= t)) / len(labels)) for t in thresholds]


def stats_lines(res):
    """The block `final_test` prints and writes to stats.txt (eff:489-493)."""
    cm = res["confmat"]
    return ["\n===================================",
            f"mAP: {res['ap']}, auroc: {res['auroc']}, acc: {res['acc']}",
            f"P@0.5: {res['precision']}, R@0.5: {res['recall']}, F1@0.5: {res['f1']}",
            f"Confmat: \n\t{cm[0][0]} | {cm[0][1]} \n\t{cm[1][0]} | {cm[1][1]}",
            "----------------------------"]


def write_stats(stats_file, res):
    with open(stats_file, "w") as f:
        for line in stats_lines(res):
            f.write(line + "\n")


def write_predictions_csv(preds_file, clips, filenames, logits, labels, ttcs):
    """predictions.csv with the columns of eff:525-533 (index, clip, filename, logits_safe, logits_risk, label, ttc),
    readable by the reference's anaysis/ scripts."""
    logits = torch.as_tensor(logits).detach().float().cpu().numpy()
    labels = torch.as_tensor(labels).detach().cpu().numpy().astype(int)
    ttcs = torch.as_tensor(ttcs).detach().cpu().numpy()
    with open(preds_file, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["", "clip", "filename", "logits_safe", "logits_risk", "label", "ttc"])
        for i in range(len(labels)):
            w.writerow([i, clips[i], filenames[i], repr(float(logits[i, 0])), repr(float(logits[i, 1])), int(labels[i]),
                        repr(float(ttcs[i]))])
