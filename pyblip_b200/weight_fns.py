"""Default weight functions defining the power BLiP maximises (same names and values as
/root/reference/pyblip/weight_fns.py)."""
import numpy as np


def inverse_size(cand_group):
    """1 / |group|"""
    return 1.0 / len(cand_group.group)


def log_inverse_size(cand_group):
    """1 / (1 + log2 |group|)"""
    return 1.0 / (1 + np.log2(len(cand_group.group)))


def inverse_radius_weight(cand_group):
    """1 / radius (continuous regions)"""
    return 1.0 / cand_group.data['radius']


def inverse_area_weight(cand_group):
    """1 / radius^2 (continuous regions)"""
    return 1.0 / (cand_group.data['radius'] ** 2)
