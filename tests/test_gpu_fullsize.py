"""BASELINE.json full sizes on the GPU, checked through size-independent properties and through
oracle comparisons on slabs the oracle finishes in seconds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _rel(a, b):
    scale = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max()) / scale


@pytest.fixture(scope="module")
def membrane2048(pkg, cuda_lib):
    from pdenlp_b200.model import GridapPDENLPModel
    spec = pkg.controlelasticmembrane(2048)
    nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False)
    yield spec, nlp
    nlp.close()


def test_membrane_n2048_sizes(membrane2048):
    spec, nlp = membrane2048
    assert nlp.meta.nvar == 16769025 + 4198401 and nlp.meta.ncon == 16769025
    assert nlp.meta.nnzj == 490266740 and nlp.meta.nnzh == 762773640
    assert nlp.pdemeta.nnzh_obj == 381386820


def test_membrane_n2048_stream_order_after_overlapped_fill(membrane2048):
    """hess_coord! launches its zero fill behind the emitting kernel with programmatic stream serialization (the
    fill may run concurrently and finishes first).  Work enqueued on the stream right after the callback must
    still see the COMPLETE result: poison the output, call, snapshot with a stream-ordered copy."""
    spec, nlp = membrane2048
    x = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, nlp.meta.nvar)).cuda()
    lam = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, nlp.meta.ncon)).cuda()
    hv = torch.empty(nlp.meta.nnzh, dtype=torch.float64, device="cuda")
    jv = torch.empty(nlp.meta.nnzj, dtype=torch.float64, device="cuda")
    ref = None
    for rep in range(4):
        hv.fill_(float("nan")); jv.fill_(float("nan"))
        nlp.hess_coord_(x, lam, hv, obj_weight=1.0)
        snap = hv.clone()                    # enqueued immediately behind the callback
        nlp.jac_coord_(x, jv)                # the next callback, then another immediate consumer
        jsum = jv.sum()
        torch.cuda.synchronize()
        assert not torch.isnan(snap).any() and torch.equal(snap, hv)
        assert torch.isfinite(jsum)
        if ref is None:
            ref = snap
        else:
            assert torch.equal(snap, ref)    # FE blocks are written without atomics: bitwise reproducible
    del snap, ref, hv, jv


def test_membrane_n2048_slabs_vs_oracle_and_properties(membrane2048, oracle_mod):
    spec, nlp = membrane2048
    nvar, ncon = nlp.meta.nvar, nlp.meta.ncon
    x = torch.from_numpy(np.random.default_rng(1998).uniform(-1, 1, nvar)).cuda()
    lam = torch.from_numpy(np.random.default_rng(1999).uniform(-1, 1, ncon)).cuda()
    jv = nlp.jac_coord(x)
    hv = nlp.hess_coord(x, lam, obj_weight=0.37)
    assert torch.isfinite(jv).all() and torch.isfinite(hv).all()
    pm = nlp.pdemeta
    ncell = spec.model.ncells
    nslab = 4096  # two mesh rows
    xh, lh = x.cpu().numpy(), lam.cpu().numpy()
    # first and last slabs: their slices sit at the head / tail of every segment
    for c0, c1, head in ((0, nslab, True), (ncell - nslab, ncell, False)):
        o = oracle_mod.Oracle(spec, (c0, c1))
        jo = o.jac_coord(xh); ho = o.hess_coord(xh, lh, 0.37)
        jy = jv[: o.nnzj_y] if head else jv[pm.nnzj_y - o.nnzj_y: pm.nnzj_y]
        ju = jv[pm.nnzj_y: pm.nnzj_y + o.nnzj_u] if head else jv[pm.nnzj_y + pm.nnzj_u - o.nnzj_u:]
        assert _rel(jy.cpu().numpy(), jo[: o.nnzj_y]) <= RTOL and _rel(ju.cpu().numpy(), jo[o.nnzj_y:]) <= RTOL
        fo = hv[: o.nnzh_fe] if head else hv[pm.nnzh_fe - o.nnzh_fe: pm.nnzh_fe]
        fc = hv[pm.nnzh_obj: pm.nnzh_obj + o.nnzh_fe] if head else hv[pm.nnzh_obj + pm.nnzh_fe - o.nnzh_fe:]
        assert _rel(fo.cpu().numpy(), ho[: o.nnzh_fe]) <= RTOL
        assert np.abs(fc.cpu().numpy() - ho[o.nnzh_fe:]).max() <= RTOL * max(np.abs(ho).max(), 1e-300)
    # a middle slab through a slab-owning handle (the multi-GPU layout): bitwise equal to the oracle-checked path
    from pdenlp_b200.model import GridapPDENLPModel
    c0 = (ncell // 2 // 32) * 32
    mid = GridapPDENLPModel(spec, device="cuda:0", cell_range=(c0, c0 + nslab), with_coef=False)
    o = oracle_mod.Oracle(spec, (c0, c0 + nslab))
    assert _rel(mid.jac_coord(x).cpu().numpy(), o.jac_coord(xh)) <= RTOL
    assert _rel(mid.hess_coord(x, lam, obj_weight=0.37).cpu().numpy()[: o.nnzh_fe], o.hess_coord(xh, lh, 0.37)[: o.nnzh_fe]) <= RTOL
    mid.close()
    # properties of this integrand: derivatives do not depend on x or lambda,
    # the residual is linear so the constraint Hessian block vanishes, obj part scales with obj_weight
    x2 = torch.from_numpy(np.random.default_rng(7).uniform(-1, 1, nvar)).cuda()
    assert torch.equal(nlp.jac_coord(x2), jv)
    assert torch.all(hv[pm.nnzh_obj:] == 0)
    h1 = nlp.hess_coord(x2, obj_weight=1.0)
    assert torch.all(h1[pm.nnzh_obj:] == 0)
    assert float((h1[: pm.nnzh_obj] * 0.37 - hv[: pm.nnzh_obj]).abs().max()) <= RTOL * float(h1.abs().max())
    # checksum of checksums (Appendix B): per cell H^f_e = w at (bubble,bubble) + w*alpha/16 on the
    # 4x4 control block, of which the lower triangle on global ids (10 entries) is stored
    w = float(np.prod(spec.model.h))
    expect = spec.model.ncells * w * (1.0 + 1e-2 * 10.0 / 16.0)
    assert abs(float(h1[: pm.nnzh_obj].sum()) - expect) <= 1e-9 * expect


def test_membrane_n2048_structure_properties(membrane2048):
    spec, nlp = membrane2048
    r, c = nlp.hess_structure()
    assert bool((r >= c).all()) and int(r.max()) <= nlp.meta.nvar and int(c.min()) >= 1
    no = nlp.pdemeta.nnzh_obj
    assert torch.equal(r[:no], r[no:]) and torch.equal(c[:no], c[no:])  # obj and cons FE blocks share the pattern
    del r, c
    r, c = nlp.jac_structure()
    assert int(r.max()) == nlp.meta.ncon and int(c.max()) == nlp.meta.nvar and int(r.min()) == 1
    # every free state dof appears as a row; y-block columns are state dofs, u-block columns control dofs
    ny = nlp.pdemeta.nnzj_y
    assert int(c[:ny].max()) <= nlp.meta.ncon and int(c[ny:].min()) > nlp.meta.ncon


def test_poissonboltzmann_n1024_lagrangian_hessian(pkg, cuda_lib, oracle_mod):
    """config 3: hess_coord! of the Lagrangian with multipliers; linear in lambda, slab vs oracle."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = pkg.poissonboltzman2d(1024)
    nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False)
    nvar, ncon = nlp.meta.nvar, nlp.meta.ncon
    assert nlp.meta.nnzh == 75374664 and nlp.meta.nnzj == 33472564
    x = torch.from_numpy(np.random.default_rng(1998).uniform(-1, 1, nvar)).cuda()
    l1 = torch.from_numpy(np.random.default_rng(1999).uniform(-1, 1, ncon)).cuda()
    l2 = torch.from_numpy(np.random.default_rng(2001).uniform(-1, 1, ncon)).cuda()
    no = nlp.pdemeta.nnzh_obj
    h1 = nlp.hess_coord(x, l1, obj_weight=0.0)
    h2 = nlp.hess_coord(x, l2, obj_weight=0.0)
    h12 = nlp.hess_coord(x, l1 + l2, obj_weight=0.0)
    assert torch.all(h1[:no] == 0)
    assert float((h1 + h2 - h12).abs().max()) <= 1e-12 * float(h12.abs().max())
    assert float(h1[no:].abs().max()) > 0  # sinh reaction term: the constraint block is not trivial
    o = oracle_mod.Oracle(spec, (0, 8192))
    ho = o.hess_coord(x.cpu().numpy(), l1.cpu().numpy(), 0.0)
    assert _rel(h1[no: no + o.nnzh_fe].cpu().numpy(), ho[o.nnzh_fe:]) <= RTOL
    nlp.close()


def test_burger1d_1M_operator_path(pkg, cuda_lib):
    """config 4: jprod!/jtprod!/hprod! on 10^6 cells; adjoint and symmetry identities."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = pkg.burger1d_param(1_000_000)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    nvar, ncon = nlp.meta.nvar, nlp.meta.ncon
    rng = np.random.default_rng(1998)
    x = torch.from_numpy(rng.uniform(-1, 1, nvar)).cuda()
    v = torch.from_numpy(rng.uniform(-1, 1, nvar)).cuda()
    v2 = torch.from_numpy(rng.uniform(-1, 1, nvar)).cuda()
    w = torch.from_numpy(rng.uniform(-1, 1, ncon)).cuda()
    Jv = nlp.jprod(x, v); Jtw = nlp.jtprod(x, w)
    a, b = float(torch.dot(w, Jv)), float(torch.dot(Jtw, v))
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b), 1.0)  # <w, J v> = <J^T w, v>
    Hv = nlp.hprod(x, w, v, obj_weight=0.5); Hv2 = nlp.hprod(x, w, v2, obj_weight=0.5)
    a, b = float(torch.dot(v2, Hv)), float(torch.dot(v, Hv2))
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b), 1.0)  # symmetry of the Lagrangian Hessian
    # coord and operator forms agree: J v from the COO values
    r, c = nlp.jac_structure(); vals = nlp.jac_coord(x)
    Jv_coo = torch.zeros(ncon, dtype=torch.float64, device="cuda").index_add_(0, r - 1, vals * v[c - 1])
    assert float((Jv_coo - Jv).abs().max()) <= 1e-12 * float(Jv.abs().max())
    nlp.close()


def test_kappa_poisson3d_128_slabs_and_identities(pkg, cuda_lib, oracle_mod):
    """config 5 (kappa + 3-D Q1 control, 128^3 cells, nq = 8): FE segments of a cell slab against the
    oracle, kappa blocks through linearity in lambda / adjoint identities, sizes from the closed forms."""
    from pdenlp_b200.model import GridapPDENLPModel
    n = 128
    spec = pkg.poissonmixed(n, dim=3, control=True)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    nvar, ncon, p = nlp.meta.nvar, nlp.meta.ncon, 2
    assert ncon == (n - 1) ** 3 and nvar == p + (n - 1) ** 3 + (n + 1) ** 3
    assert nlp.pdemeta.nnzj_k == ncon * p and nlp.pdemeta.nnzh_k_obj == 3 and nlp.pdemeta.nnzh_k_con == 3 + (nvar - p) * p
    rng = np.random.default_rng(1998)
    xh = rng.uniform(-1, 1, nvar); xh[:p] = 1 + 0.1 * rng.uniform(-1, 1, p)
    lh = rng.uniform(-1, 1, ncon)
    x, lam = torch.from_numpy(xh).cuda(), torch.from_numpy(lh).cuda()
    jv = nlp.jac_coord(x); hv = nlp.hess_coord(x, lam, obj_weight=0.37)
    # a slab of 2 mesh planes from the middle against the oracle (FE segments are cell-major)
    c0 = (n // 2) * n * n; c1 = c0 + 2 * n * n
    o = oracle_mod.Oracle(spec, (c0, c1))
    full = nlp.handle  # noqa: F841  (offsets of the slab inside the global segments come from a second handle)
    head = GridapPDENLPModel(spec, device="cuda:0", cell_range=(0, c0), with_coef=False)
    oy, ou, oh = head.pdemeta.nnzj_y, head.pdemeta.nnzj_u, head.pdemeta.nnzh_fe
    head.close()
    jo = o.jac_coord(xh); ho = o.hess_coord(xh, lh, 0.37)
    pm = nlp.pdemeta
    a = pm.nnzj_k + oy
    assert _rel(jv[a: a + o.nnzj_y].cpu().numpy(), jo[o.nnzj_k: o.nnzj_k + o.nnzj_y]) <= RTOL
    a = pm.nnzj_k + pm.nnzj_y + ou
    assert _rel(jv[a: a + o.nnzj_u].cpu().numpy(), jo[o.nnzj_k + o.nnzj_y:]) <= RTOL
    a = pm.nnzh_k_obj + oh
    assert _rel(hv[a: a + o.nnzh_fe].cpu().numpy(), ho[o.nnzh_k_obj: o.nnzh_k_obj + o.nnzh_fe]) <= RTOL
    a = pm.nnzh_obj + pm.nnzh_k_con + oh
    assert torch.all(hv[a: a + o.nnzh_fe] == 0)  # k1*grad(v).grad(y) - v*f*k2 - v*u: affine in (y, u)
    assert np.all(ho[o.nnzh_obj + o.nnzh_k_con:] == 0)
    # kappa blocks: <w, J v> = <J^T w, v> with v exciting the kappa columns, symmetry of H
    v = torch.from_numpy(rng.uniform(-1, 1, nvar)).cuda(); v2 = torch.from_numpy(rng.uniform(-1, 1, nvar)).cuda()
    w = torch.from_numpy(rng.uniform(-1, 1, ncon)).cuda()
    a_, b_ = float(torch.dot(w, nlp.jprod(x, v))), float(torch.dot(nlp.jtprod(x, w), v))
    assert abs(a_ - b_) <= 1e-11 * max(abs(a_), abs(b_), 1.0)
    Hv = nlp.hprod(x, lam, v, obj_weight=0.37); Hv2 = nlp.hprod(x, lam, v2, obj_weight=0.37)
    a_, b_ = float(torch.dot(v2, Hv)), float(torch.dot(v, Hv2))
    assert abs(a_ - b_) <= 1e-11 * max(abs(a_), abs(b_), 1.0)
    # coord form == operator form, kappa blocks included
    r, c = nlp.hess_structure()
    d = r != c
    Hv_coo = torch.zeros(nvar, dtype=torch.float64, device="cuda").index_add_(0, r - 1, hv * v[c - 1])
    Hv_coo.index_add_(0, c[d] - 1, hv[d] * v[r[d] - 1])
    assert float((Hv_coo - Hv).abs().max()) <= 1e-11 * float(Hv.abs().max())
    nlp.close()


def test_burger1d_1M_coord_slabs_vs_oracle(pkg, cuda_lib, oracle_mod):
    """config 4: jac_coord!/hess_coord! on 10^6 cells; a slab from the middle (FE segments are cell-major, so the
    slab's runs sit at the offsets given by a head model) and the kappa blocks against the oracle of the full mesh."""
    from pdenlp_b200.model import GridapPDENLPModel
    n = 1_000_000
    spec = pkg.burger1d_param(n)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    rng = np.random.default_rng(1998)
    xh = rng.uniform(-1, 1, nlp.meta.nvar); xh[0] = 1.05
    lh = rng.uniform(-1, 1, nlp.meta.ncon)
    x, lam = torch.from_numpy(xh).cuda(), torch.from_numpy(lh).cuda()
    jv = nlp.jac_coord(x); hv = nlp.hess_coord(x, lam, obj_weight=0.37)
    c0, c1 = 400_000, 465_536
    head = GridapPDENLPModel(spec, device="cuda:0", cell_range=(0, c0), with_coef=False)
    oy, ou, oh = head.pdemeta.nnzj_y, head.pdemeta.nnzj_u, head.pdemeta.nnzh_fe
    head.close()
    o = oracle_mod.Oracle(spec, (c0, c1))
    jo = o.jac_coord(xh); ho = o.hess_coord(xh, lh, 0.37)
    pm = nlp.pdemeta
    a = pm.nnzj_k + oy
    assert _rel(jv[a: a + o.nnzj_y].cpu().numpy(), jo[o.nnzj_k: o.nnzj_k + o.nnzj_y]) <= RTOL
    a = pm.nnzj_k + pm.nnzj_y + ou
    assert _rel(jv[a: a + o.nnzj_u].cpu().numpy(), jo[o.nnzj_k + o.nnzj_y:]) <= RTOL
    a = pm.nnzh_k_obj + oh
    assert _rel(hv[a: a + o.nnzh_fe].cpu().numpy(), ho[o.nnzh_k_obj: o.nnzh_k_obj + o.nnzh_fe]) <= RTOL
    a = pm.nnzh_obj + pm.nnzh_k_con + oh
    assert _rel(hv[a: a + o.nnzh_fe].cpu().numpy(), ho[o.nnzh_obj + o.nnzh_k_con:]) <= RTOL
    # dense kappa blocks (atomics over all cells) against the full-mesh oracle
    of = oracle_mod.Oracle(spec)
    jf = of.jac_coord(xh); hf = of.hess_coord(xh, lh, 0.37)
    assert _rel(jv[: pm.nnzj_k].cpu().numpy(), jf[: of.nnzj_k]) <= 1e-11
    assert _rel(hv[: pm.nnzh_k_obj].cpu().numpy(), hf[: of.nnzh_k_obj]) <= 1e-11
    b = pm.nnzh_obj
    assert _rel(hv[b: b + pm.nnzh_k_con].cpu().numpy(), hf[of.nnzh_obj: of.nnzh_obj + of.nnzh_k_con]) <= 1e-11
    nlp.close()


@pytest.mark.parametrize("case", ["membrane2048", "kappa_poisson3d_128"])
def test_template_path_equals_cell_by_cell_evaluation_over_whole_outputs(pkg, cuda_lib, case):
    """Full-size property (BASELINE configs[1] and configs[4]): the template path (chunks copied from the image of one
    template cell + lane = entry emission of the boundary cells) against the generic kernels that evaluate every
    cell (PDENLP_FLAG_NO_TEMPLATES), over the WHOLE jac_coord! / hess_coord! value streams — every chunk offset, every
    boundary cell, head to tail.  Same fma chains on both sides: the FE segments must agree bit for bit; the dense
    kappa blocks are atomic sums (round-off)."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = pkg.controlelasticmembrane(2048) if case == "membrane2048" else pkg.poissonmixed(128, dim=3, control=True)
    a = GridapPDENLPModel(spec, device="cuda:0", with_coef=(spec.nparam > 0))
    b = GridapPDENLPModel(spec, device="cuda:0", with_coef=(spec.nparam > 0), flat=a.flat, templates=False)
    try:
        I, pm = a.handle.info, a.pdemeta
        rng = np.random.default_rng(77)
        xh = rng.uniform(-1, 1, I.nvar)
        if spec.nparam:
            xh[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
        x = torch.from_numpy(xh).cuda()
        lam = torch.from_numpy(rng.uniform(-1, 1, I.ncon)).cuda()
        nk = int(I.nnzj_k)
        ja, jb = a.jac_coord(x), b.jac_coord(x)
        assert torch.equal(ja[nk:], jb[nk:])
        if nk:
            assert float((ja[:nk] - jb[:nk]).abs().max()) <= 1e-12 * float(jb[:nk].abs().max())
        del ja, jb
        for ow in (1.0, 0.37):
            ha, hb = a.hess_coord(x, lam, obj_weight=ow), b.hess_coord(x, lam, obj_weight=ow)
            ko, kc = int(I.nnzh_k_obj), int(I.nnzh_k_con)
            fo0, fo1 = ko, ko + int(I.nnzh_fe_obj)
            assert torch.equal(ha[fo0:fo1], hb[fo0:fo1])            # objective FE block
            assert torch.equal(ha[fo1 + kc:], hb[fo1 + kc:])        # constraint FE block
            for lo, hi in ((0, ko), (fo1, fo1 + kc)):
                if hi > lo:
                    assert float((ha[lo:hi] - hb[lo:hi]).abs().max()) <= 1e-12 * max(float(hb[lo:hi].abs().max()), 1e-300)
            del ha, hb
    finally:
        a.close(); b.close()


TEMPLATE_FAMILIES = {
    # every instantiated element with a template path, at a size where it is active (>= 4096 tiles of 32 cells)
    "membrane_q1q1_nq1": lambda pkg: pkg.controlelasticmembrane(384, state_order=1),
    "membrane_q1q1_nq4": lambda pkg: pkg.controlelasticmembrane(384, state_order=1, degree=2),
    "membrane_q2q1_nq1": lambda pkg: pkg.controlelasticmembrane(384),
    "membrane_q2q1_nq4": lambda pkg: pkg.controlelasticmembrane(384, degree=2),
    "membrane3d_q1q1_nq1": lambda pkg: pkg.controlelasticmembrane(56, state_order=1, dim=3),
    "membrane3d_q1q1_nq8": lambda pkg: pkg.controlelasticmembrane(56, state_order=1, degree=2, dim=3),
    "kappa_poisson2d": lambda pkg: pkg.poissonmixed(384, 2, False),
    "kappa_poisson2d_control": lambda pkg: pkg.poissonmixed(384, 2, True),
    "kappa_poisson3d_control": lambda pkg: pkg.poissonmixed(56, 3, True),
    "poissonparam": lambda pkg: pkg.poissonparam(384),
    "penalizedpoisson": lambda pkg: pkg.penalizedpoisson(384),
    "basicunconstrained": lambda pkg: pkg.basicunconstrained(384),
}


@pytest.mark.parametrize("family", sorted(TEMPLATE_FAMILIES))
def test_template_path_equals_cell_by_cell_evaluation_for_every_element_family(pkg, cuda_lib, family):
    """The cell-granular template path against the generic kernels (PDENLP_FLAG_NO_TEMPLATES) over whole outputs, for
    every (integrand, element, quadrature) instance that has one: FE segments bit for bit (same fma chains), dense
    kappa blocks to round-off.  The small parity cases of test_gpu_parity.py stay below the 4096-tile threshold and
    never take this path."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = TEMPLATE_FAMILIES[family](pkg)
    a = GridapPDENLPModel(spec, device="cuda:0")
    b = GridapPDENLPModel(spec, device="cuda:0", flat=a.flat, templates=False)
    try:
        I = a.handle.info
        assert a.flat.ncells >= 4096 * 32
        sa, sb = a.handle.template_stats(), b.handle.template_stats()
        assert sa["chunks"] > 0 and sa["template_cells"] + sa["generic_cells"] == a.flat.ncells, sa  # the path is taken
        assert sa["template_cells"] > 0.85 * a.flat.ncells and sb["chunks"] == 0, (sa, sb)  # (3-D n = 56: 54^3 / 56^3 = 0.897)
        rng = np.random.default_rng(5)
        xh = rng.uniform(-1, 1, I.nvar)
        if spec.nparam:
            xh[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
        x = torch.from_numpy(xh).cuda()
        ncon = a.meta.ncon  # (0 for the unconstrained problems)
        lam = torch.from_numpy(rng.uniform(-1, 1, ncon)).cuda()

        def same(u, v, dense):  # dense = [(lo, hi)] ranges of atomic sums inside the stream
            mask = torch.ones(u.numel(), dtype=torch.bool, device=u.device)
            for lo, hi in dense:
                mask[lo:hi] = False
                if hi > lo:
                    assert float((u[lo:hi] - v[lo:hi]).abs().max()) <= 1e-12 * max(float(v[lo:hi].abs().max()), 1e-300)
            assert torch.equal(u[mask], v[mask])

        if ncon > 0:
            same(a.jac_coord(x), b.jac_coord(x), [(0, int(I.nnzj_k))])
        ko, kc, fo = int(I.nnzh_k_obj), int(I.nnzh_k_con), int(I.nnzh_fe_obj)
        for ow in (1.0, 0.37, 0.0):
            if ncon > 0:
                same(a.hess_coord(x, lam, obj_weight=ow), b.hess_coord(x, lam, obj_weight=ow), [(0, ko), (ko + fo, ko + fo + kc)])
            same(a.hess_coord(x, obj_weight=ow), b.hess_coord(x, obj_weight=ow), [(0, ko)])
    finally:
        a.close(); b.close()


@pytest.mark.parametrize("family", ["membrane_q2q1_nq1", "kappa_poisson2d_control"])
def test_template_path_through_the_host_memory_space(pkg, cuda_lib, family):
    """PDENLP_MEM_HOST handles (the Julia stub's default: host vectors staged by the library) at a template-active size:
    same numbers as the device-memspace model, including the structurally zero tail that host threads fill."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = TEMPLATE_FAMILIES[family](pkg)
    d = GridapPDENLPModel(spec, device="cuda:0")
    h = GridapPDENLPModel(spec, device="cuda:0", flat=d.flat, memspace="host")
    try:
        assert h.handle.template_stats()["chunks"] > 0
        rng = np.random.default_rng(9)
        x = rng.uniform(-1, 1, d.meta.nvar)
        if spec.nparam:
            x[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
        lam = rng.uniform(-1, 1, d.meta.ncon)
        xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lam).cuda()
        I = d.handle.info
        nk = int(I.nnzj_k)
        for rep in range(2):  # (second pass: cached images)
            jh, jd = h.jac_coord(x), d.jac_coord(xd).cpu().numpy()
            assert np.array_equal(jh[nk:], jd[nk:]) and np.allclose(jh[:nk], jd[:nk], rtol=1e-12, atol=1e-14)
            hh, hd = h.hess_coord(x, lam, obj_weight=0.37), d.hess_coord(xd, ld, obj_weight=0.37).cpu().numpy()
            ko, kc, fo = int(I.nnzh_k_obj), int(I.nnzh_k_con), int(I.nnzh_fe_obj)
            assert np.array_equal(hh[ko: ko + fo], hd[ko: ko + fo]) and np.array_equal(hh[ko + fo + kc:], hd[ko + fo + kc:])
            assert np.allclose(hh, hd, rtol=1e-12, atol=1e-14 * np.abs(hd).max())
    finally:
        d.close(); h.close()


@pytest.mark.parametrize("family", ["membrane_q2q1_nq1", "kappa_poisson3d_control"])
@pytest.mark.parametrize("shift", [0, 1, 3])
def test_template_path_writes_exactly_its_output_range(pkg, cuda_lib, family, shift):
    """Guard bands: the value arrays are views into larger NaN-filled buffers, starting `shift` doubles into them (an odd
    shift makes every 16 B-aligned store of the chunk copy start one element late: head / tail paths).  Nothing outside
    [0, nnz) may be written, nothing inside may stay unwritten, and the values do not depend on the alignment."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = TEMPLATE_FAMILIES[family](pkg)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    try:
        assert nlp.handle.template_stats()["chunks"] > 0
        rng = np.random.default_rng(21)
        xh = rng.uniform(-1, 1, nlp.meta.nvar)
        if spec.nparam:
            xh[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
        x = torch.from_numpy(xh).cuda()
        lam = torch.from_numpy(rng.uniform(-1, 1, nlp.meta.ncon)).cuda()
        pad = 4096
        ref_j, ref_h = nlp.jac_coord(x), nlp.hess_coord(x, lam, obj_weight=0.37)
        I = nlp.handle.info
        for nnz, call, ref, dense in ((nlp.meta.nnzj, lambda v: nlp.jac_coord_(x, v), ref_j, [(0, int(I.nnzj_k))]),
                                      (nlp.meta.nnzh, lambda v: nlp.hess_coord_(x, lam, v, obj_weight=0.37), ref_h,
                                       [(0, int(I.nnzh_k_obj)), (int(I.nnzh_obj), int(I.nnzh_obj) + int(I.nnzh_k_con))])):
            buf = torch.full((nnz + 2 * pad + 8,), float("nan"), dtype=torch.float64, device="cuda")
            view = buf[pad + shift: pad + shift + nnz]
            call(view)
            torch.cuda.synchronize()
            assert torch.isnan(buf[: pad + shift]).all() and torch.isnan(buf[pad + shift + nnz:]).all(), "wrote outside its range"
            assert not torch.isnan(view).any(), "left entries unwritten"
            mask = torch.ones(nnz, dtype=torch.bool, device="cuda")
            for lo, hi in dense:
                mask[lo:hi] = False
            assert torch.equal(view[mask], ref[mask])
            assert torch.allclose(view, ref, rtol=1e-12, atol=1e-14 * float(ref.abs().max()))
            del buf
    finally:
        nlp.close()
