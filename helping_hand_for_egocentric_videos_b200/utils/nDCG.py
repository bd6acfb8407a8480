"""Mirror of the reference's ``utils/nDCG.py`` (:3-150): same five functions, same arguments and return values, with
the ranking and the discounted sums on the retrieval kernel of libhh_b200.so (float64, numpy's summation order)."""
import numpy as np

from .. import ops


def calculate_DCG(similarity_matrix, relevancy_matrix, k_counts):
    """DCG per query of the first modality (n1 vector); k_counts is the n1 x n2 mask of ranks that count."""
    return ops.retrieval_rows(similarity_matrix, relevancy_matrix, mode=1, kcounts=k_counts)


def calculate_k_counts(relevancy_matrix):
    """(np.sort(rel)[:, ::-1] > 0).astype(int) of the reference (:75): positives sort first, so rank k counts iff
    k < number of positive relevancies of the row -- no sort needed."""
    rel = np.asarray(relevancy_matrix)
    npos = (rel > 0).sum(axis=1)
    return (np.arange(rel.shape[1])[None, :] < npos[:, None]).astype(int)


def calculate_IDCG(relevancy_matrix, k_counts):
    return calculate_DCG(relevancy_matrix, relevancy_matrix, k_counts)


def calculate_nDCG(similarity_matrix, relevancy_matrix, k_counts=None, IDCG=None, reduction='mean'):
    if k_counts is None:
        k_counts = calculate_k_counts(relevancy_matrix)
    DCG = calculate_DCG(similarity_matrix, relevancy_matrix, k_counts)
    if IDCG is None:
        IDCG = calculate_IDCG(relevancy_matrix, k_counts)
    if reduction == 'mean':
        return np.mean(DCG / IDCG)
    elif reduction is None:
        return DCG / IDCG
