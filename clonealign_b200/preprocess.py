"""`preprocess_for_clonealign()` host mirror (R/preprocess.R:93-147): one-shot filtering before the hot path."""
from __future__ import annotations

import numpy as np


def get_outlying_genes(Y, nmads):
    """R/preprocess.R:58-62 — stats::mad uses the 1.4826 consistency constant."""
    gene_means = np.asarray(Y, dtype=np.float64).mean(axis=0)
    md = 1.4826 * np.median(np.abs(gene_means - np.median(gene_means)))
    return gene_means > gene_means.mean() + nmads * md


def preprocess_for_clonealign(gene_expression_data, copy_number_data, min_counts_per_gene=20, min_counts_per_cell=100,
                              remove_outlying_genes=True, nmads=10, max_copy_number=6,
                              remove_genes_same_copy_number=True, device=None):
    """Returns dict(gene_expression_data, copy_number_data, retained_cells, retained_genes) (0-based indices).
    `device=<CUDA ordinal>`: the two passes over the N x G matrix that the filters need -- colSums(Y) (:118, :59) and
    rowSums over the retained genes (:138) -- run as reductions on the matrix resident in HBM (ca_core_data_stats,
    ca_core_data_masked_rowsums) instead of on the host; the gene filters themselves only touch G-vectors."""
    if device is not None:
        return _preprocess_on_device(gene_expression_data, copy_number_data, min_counts_per_gene, min_counts_per_cell,
                                     remove_outlying_genes, nmads, max_copy_number, remove_genes_same_copy_number, device)
    Y = np.asarray(gene_expression_data)
    L = np.asarray(copy_number_data, dtype=np.float64)
    if L.shape[0] != Y.shape[1]:
        raise ValueError("copy_number_data must have same number of genes (rows) as gene_expression_data")
    genes = np.arange(Y.shape[1])
    cells = np.arange(Y.shape[0])

    def keep_genes(mask):
        nonlocal Y, L, genes
        Y, L, genes = Y[:, mask], L[mask], genes[mask]

    keep_genes(~(L.max(axis=1) > max_copy_number))                 # :114-116
    keep_genes(Y.sum(axis=0) > min_counts_per_gene)                # :118-120
    if remove_outlying_genes:
        keep_genes(~get_outlying_genes(Y, nmads))                  # :123-128
    if remove_genes_same_copy_number:
        keep_genes(~(L.var(axis=1, ddof=1) == 0))                  # :131-135
    cwc = Y.sum(axis=1) > min_counts_per_cell                      # :138-139
    Y, cells = Y[cwc], cells[cwc]
    return dict(gene_expression_data=Y, copy_number_data=L, retained_cells=cells, retained_genes=genes)


def _preprocess_on_device(Y_in, copy_number_data, min_counts_per_gene, min_counts_per_cell, remove_outlying_genes, nmads,
                          max_copy_number, remove_genes_same_copy_number, device):
    """The same filters with the matrix passes on the GPU.  The column sums do not depend on which genes have been dropped
    so far (columns are independent), and gene_means = colSums / N (:59) over the genes that are still in: every gene
    filter is a G-vector operation on the host once colSums(Y) is known."""
    from .session import DeviceData
    sparse = hasattr(Y_in, "tocsr") and hasattr(Y_in, "nnz")
    Y = Y_in.tocsr() if sparse else np.asarray(Y_in)
    L = np.asarray(copy_number_data, dtype=np.float64)
    N, G = Y.shape
    if L.shape[0] != G:
        raise ValueError("copy_number_data must have same number of genes (rows) as gene_expression_data")
    with DeviceData(Y, np.where(L > 0, L, 1.0), device=device) as dd:     # the copy-number argument is not used by the sums
        colsum = dd.stats()["colsum"]
        keep = ~(L.max(axis=1) > max_copy_number)                          # :114-116
        keep &= colsum > min_counts_per_gene                               # :118-120
        if remove_outlying_genes:                                          # :123-128, on the genes still in
            gm = colsum[keep] / N
            md = 1.4826 * np.median(np.abs(gm - np.median(gm)))
            out = np.zeros(G, dtype=bool)
            out[np.nonzero(keep)[0]] = gm > gm.mean() + nmads * md
            keep &= ~out
        if remove_genes_same_copy_number:                                  # :131-135
            keep &= ~(L.var(axis=1, ddof=1) == 0)
        rows = dd.masked_rowsums(keep)                                     # :138
    cwc = rows > min_counts_per_cell
    genes, cells = np.nonzero(keep)[0], np.nonzero(cwc)[0]
    Yk = Y[cells][:, genes]
    return dict(gene_expression_data=Yk, copy_number_data=L[genes], retained_cells=cells, retained_genes=genes)
