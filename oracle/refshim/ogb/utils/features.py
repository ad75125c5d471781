"""ogb 1.2.4 feature-dimension constants (the only thing the reference's model code needs)."""


def get_atom_feature_dims():
    return [119, 4, 12, 12, 10, 6, 6, 2, 2]


def get_bond_feature_dims():
    return [5, 6, 2]
