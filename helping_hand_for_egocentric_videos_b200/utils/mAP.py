"""Mirror of the reference's ``utils/mAP.py`` (calculate_mAP, :4-44) on the retrieval kernel of libhh_b200.so: the
N x M ranking, the gather of the relevancies and the per-query average precision run on the device in float64; only the
N per-query values come back for numpy's mean."""
import numpy as np

from .. import ops


def calculate_mAP(sim_mat, relevancy_matrix):
    avg_precision = ops.retrieval_rows(sim_mat, relevancy_matrix, mode=0)
    return np.mean(avg_precision)
