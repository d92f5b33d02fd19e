"""Shared by the r02 probes: the example-1 system of a level (condensed K2, rhs = K2c x*)."""
import numpy as np, torch
import fem_forth_perturb_b200 as F

def system(level, eps=1e-2):
    gm = F.MeshTri().refine(level)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(eps ** 2 * F.a_load + F.b_load, basis)
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    g = torch.Generator(device='cuda').manual_seed(1234)
    x = torch.rand(K.shape[0], dtype=torch.float64, device='cuda', generator=g) * 2 - 1
    x[torch.from_numpy(D).cuda()] = 0
    f = K.dot(x)
    Kc, bc, _, I = F.condense(K, f, D=D)
    return Kc, bc

