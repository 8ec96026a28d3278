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
                              remove_genes_same_copy_number=True):
    """Returns dict(gene_expression_data, copy_number_data, retained_cells, retained_genes) (0-based indices)."""
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
