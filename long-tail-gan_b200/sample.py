"""Sampler plugin with the reference's scalar signature (Codes/sample.py:40-67) for callers that hold one user's
candidate probabilities on the host. The training path does not come through here: phase A samples every user of the
batch on the device (ltg_sample_pairs, engine.GanEngine.phase_a). This shim exists so code written against the
reference's sample.py keeps working; it draws the same distribution (successive draws without replacement)."""
import numpy as np


def sample_from_generator_new(elements, probabilities_li, to_sample, num_elements):
    """Returns (binary mask float64 [num_elements], sampled ids) like sample.py:40-67."""
    sampled_li_bin = np.zeros([num_elements], dtype=float)
    p = np.asarray(probabilities_li, dtype=np.float64)
    if p.sum() != 0.0:
        p = p / (1.0 * p.sum())
    else:
        p = np.full(len(elements), 1.0 / max(1, len(elements)))
    elements = np.asarray(elements)
    to_sample = int(min(to_sample, np.count_nonzero(p)))  # sample.py:51-61 shrinks the draw until np.random.choice accepts it
    if to_sample <= 0:
        return sampled_li_bin, np.asarray([], dtype=elements.dtype)
    # Gumbel-top-k == Plackett-Luce == np.random.choice(..., replace=False, p=p) in distribution
    with np.errstate(divide="ignore"):
        keys = np.log(p) + np.random.gumbel(size=len(p))
    sampled = elements[np.argpartition(-keys, to_sample - 1)[:to_sample]]
    sampled_li_bin[sampled] = 1
    return sampled_li_bin, np.asarray(sampled)
