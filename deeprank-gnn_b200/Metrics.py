"""Classification / regression metrics of an epoch, attribute-compatible with the reference's
``Metrics`` (``deeprank_gnn/Metrics.py:69-260``): CPU post-processing on Python lists, outside
the hot path."""
import numpy as np


def get_binary(values, threshold, target):
    """1 = 'good': above the threshold for fnat / bin_class, below it otherwise (Metrics.py:10-32)."""
    if target in ('fnat', 'bin_class'):
        return [1 if x > threshold else 0 for x in values]
    return [1 if x < threshold else 0 for x in values]


def get_comparison(prediction, ground_truth, binary=True, classes=(0, 1)):
    classes = list(classes)
    idx = {c: i for i, c in enumerate(classes)}
    cm = np.zeros((len(classes), len(classes)), dtype=np.int64)
    for p, t in zip(prediction, ground_truth):
        if p in idx and t in idx:
            cm[idx[t], idx[p]] += 1
    fp = cm.sum(axis=0) - np.diag(cm)
    fn = cm.sum(axis=1) - np.diag(cm)
    tp = np.diag(cm)
    tn = cm.sum() - (fp + fn + tp)
    if binary:
        return fp[1], fn[1], tp[1], tn[1]
    return fp, fn, tp, tn


def _ratio(a, b):
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.divide(a, b)
    return r


class Metrics(object):
    def __init__(self, prediction, y, target, threshold=4, binary=True):
        self.prediction, self.y, self.binary, self.target, self.threshold = prediction, y, binary, target, threshold
        if binary:
            pb, yb = get_binary(prediction, threshold, target), get_binary(y, threshold, target)
            fp, fn, tp, tn = get_comparison(pb, yb, True, (0, 1))
        else:
            if target == 'capri_class':
                classes = [1, 2, 3, 4, 5]
            elif target == 'bin_class':
                classes = [0, 1]
            else:
                raise ValueError('target must be capri_class on bin_class')
            fp, fn, tp, tn = get_comparison(prediction, y, False, classes)
        self.sensitivity = _ratio(tp, tp + fn)
        self.specificity = _ratio(tn, tn + fp)
        self.precision = _ratio(tp, tp + fp)
        self.NPV = _ratio(tn, tn + fn)
        self.FPR = _ratio(fp, fp + tn)
        self.FNR = _ratio(fn, tp + fn)
        self.FDR = _ratio(fp, tp + fp)
        self.accuracy = _ratio(tp + tn, tp + fp + fn + tn)
        # regression metrics: only for the continuous targets, None otherwise (Metrics.py:176-215); the attribute
        # names (including the reference's ``mean_abolute_error`` / ``median_squared_log_error``) are kept
        self.explained_variance = self.max_error = self.mean_abolute_error = self.mean_squared_error = None
        self.root_mean_squared_error = self.mean_squared_log_error = self.median_squared_log_error = None
        self.r2_score = self.mean_absolute_error = None
        if target in ('fnat', 'irmsd', 'lrmsd'):
            p, t = np.asarray(prediction, dtype=np.float64), np.asarray(y, dtype=np.float64)
            if p.size == 0 or p.shape != t.shape:
                raise ValueError('Metrics: predictions and targets must be non-empty lists of the same length')
            err = t - p
            self.max_error = float(np.abs(err).max())
            self.mean_absolute_error = float(np.abs(err).mean())
            self.mean_squared_error = float((err ** 2).mean())
            self.root_mean_squared_error = float(np.sqrt((err ** 2).mean()))
            var = float(t.var())
            # sklearn conventions for a constant target: 1.0 for a perfect fit, 0.0 otherwise
            self.explained_variance = float(1 - err.var() / var) if var > 0 else (1.0 if float(err.var()) == 0 else 0.0)
            ss_res, ss_tot = float((err ** 2).sum()), float(((t - t.mean()) ** 2).sum())
            self.r2_score = float(1 - ss_res / ss_tot) if ss_tot > 0 else (1.0 if ss_res == 0 else 0.0)
            if (p < 0).any() or (t < 0).any():
                print('WARNING: Mean Squared Logarithmic Error cannot be used when targets contain negative values.')
            else:
                self.mean_squared_log_error = float(((np.log1p(t) - np.log1p(p)) ** 2).mean())
            self.median_squared_log_error = float(np.median(np.abs(err)))      # (median ABSOLUTE error, Metrics.py:210)

    def format_score(self):
        """Ranks of the predictions (best first: descending for fnat / bin_class, ascending otherwise) and the binary
        ground truth (Metrics.py:218-239)."""
        idx = np.argsort(self.prediction)
        if self.target in ('fnat', 'bin_class'):
            idx = idx[::-1]
        return idx, np.array(get_binary(self.y, self.threshold, self.target))

    def hitrate(self):
        idx, gt = self.format_score()
        return np.cumsum(gt[idx])

    def auc(self):
        """``roc_auc_score(ground_truth, idx)`` exactly as the reference computes it (Metrics.py:251-260: the score it
        passes is the rank INDEX array, not the predictions - kept for attribute-level parity)."""
        from sklearn.metrics import roc_auc_score
        idx, gt = self.format_score()
        return roc_auc_score(gt, idx)
