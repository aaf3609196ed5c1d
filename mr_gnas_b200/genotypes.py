"""Genotype strings are API: ``Genotype(alpha_cell=[(op, center, pre), ...], concat_node=[...],
score_func='sf_DisMult')`` is ``eval``'d by the scripts (train/mr_lp_train.py:110,
train/mr_nc_train.py:75).  The reference's namedtuple (configs/genotypes.py:3) has three
mandatory fields although the NC default string gives only two, so ``score_func`` defaults to
None here."""
from collections import namedtuple

Genotype = namedtuple('Genotype', 'alpha_cell concat_node score_func', defaults=(None,))
