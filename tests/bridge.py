"""Test-side bridge: turn a host ``Solver`` description into the oracle's array dict.

Only tests / smoke / bench's cpu_baseline use this (the oracle is never on the product path).
Dense operator matrices come from ``LinearMap.to_dense()`` so the oracle exercises the dense
(BLASAlgorithm-like) path while the CUDA kernels use sum-factorised / sparse operators.
"""
import numpy as np

from sse_b200.linear_maps import IdentityMap, SelectionMap
from sse_b200.reference_approximation import NoMapping


def oracle_problem(solver):
    sd = solver.spatial_discretization
    ra = sd.reference_approximation
    gf = sd.geometric_factors
    prob = dict(
        d=ra.dim, N_p=ra.N_p, N_q=ra.N_q, N_f=ra.N_f, N_c=solver.law_desc["N_c"], N_e=sd.N_e,
        num_faces=ra.element_type.num_faces,
        V=ra.V.to_dense(), R=ra.R.to_dense(), D=[D.to_dense() for D in ra.D],
        W=ra.W, B=ra.B,
        V_is_identity=isinstance(ra.V, IdentityMap),
        R_is_selection=isinstance(ra.R, SelectionMap),
        J_q=gf.J_q, Lambda_q=gf.Lambda_q, J_f=gf.J_f, nJf=gf.nJf, n_ref=gf.n_ref,
        mapP=sd.mesh.mapP, law=solver.law_desc, form=solver.form_desc,
        mass_solver=solver.mass_kind, Minv=solver.Minv,
    )
    if not isinstance(ra.reference_mapping, NoMapping):
        prob["Lambda_ref"] = ra.reference_mapping.Lambda_ref
        prob["J_ref"] = ra.reference_mapping.J_ref
    return prob
