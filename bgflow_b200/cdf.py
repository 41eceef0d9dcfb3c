"""IC-domain maps: ``CDFTransform`` and the marginals the builder installs on internal coordinates.

Mirrors bgflow/nn/flow/cdf.py:13-125 (``CDFTransform``, ``DistributionTransferFlow``,
``ConstrainGaussianFlow``), bgflow/distribution/normal.py:95-227 (``TruncatedNormalDistribution``)
and bgflow/distribution/distributions.py:71-97 (``SloppyUniform``): same constructors, same
direction convention (``_forward`` = cdf, ``_inverse`` = icdf; the builder wraps the transform in
``InverseFlow``, generator_builder.py:451), same ``eps`` clamps.

For truncated-normal / normal / uniform marginals every call is ONE launch of ``bgx_cdf_map``
(csrc/bgx_cdf.cu) that also adds the log-det onto the running ``dlogp``; ``MultiCDFFlow`` maps
several tensors of the flow state in the same launch and ``fuse_domain_maps`` folds the maps of
bonds / angles / torsions into the IC -> Cartesian kernel (``bgx_ic_to_xyz_mapped``).  Other
distributions (anything with ``cdf / icdf / log_prob``) and learnable marginals under autograd run
the reference's op sequence on the device with torch.
"""

import math

import torch

from . import _lib, engine
from .flows import Flow, InverseFlow, SequentialFlow, SetConstantFlow, WrapFlow

__all__ = ["CDFTransform", "DistributionTransferFlow", "ConstrainGaussianFlow", "TruncatedNormalDistribution",
           "SloppyUniform", "MultiCDFFlow", "MappedICTail", "fuse_domain_maps", "marginal_columns"]


class TruncatedNormalDistribution(torch.nn.Module):
    """Normal(mu, sigma) restricted to [lower_bound, upper_bound], independent per dimension
    (normal.py:95-227): ``sample``, ``energy``, ``cdf``, ``icdf``, ``log_prob``."""

    def __init__(self, mu, sigma=torch.tensor(1.0), lower_bound=torch.tensor(0.0),
                 upper_bound=torch.tensor(math.inf), assert_range=True, sampling_method="icdf",
                 is_learnable=False):
        super().__init__()
        for t in (mu, sigma, lower_bound, upper_bound):
            assert type(t) is torch.Tensor
            assert t.shape in (torch.Size([]), (1,), mu.shape)
        self._dim = mu.shape
        self.event_shapes = [torch.Size(mu.shape)]
        if is_learnable:
            self._mu = torch.nn.Parameter(mu)
            self._logsigma = torch.nn.Parameter(torch.log(sigma.to(mu)))
        else:
            self.register_buffer("_mu", mu)
            self.register_buffer("_logsigma", torch.log(sigma.to(mu)))
        self.register_buffer("_upper_bound", upper_bound.to(mu))
        self.register_buffer("_lower_bound", lower_bound.to(mu))
        self.assert_range = assert_range
        if sampling_method not in ("icdf", "rejection"):
            raise ValueError(f'Unknown sampling method "{sampling_method}"')
        self._sampling_method = sampling_method
        alpha = (self._lower_bound - self._mu) / self._sigma
        beta = (self._upper_bound - self._mu) / self._sigma
        self.register_buffer("_cdf_lower_bound", self._std_cdf(alpha.detach()))
        self.register_buffer("_cdf_upper_bound", self._std_cdf(beta.detach()))

    @staticmethod
    def _std_cdf(z):
        return 0.5 * (1 + torch.erf(z / math.sqrt(2)))

    @staticmethod
    def _std_icdf(p):
        return torch.erfinv(2 * p - 1) * math.sqrt(2)

    @property
    def _sigma(self):
        return torch.exp(self._logsigma)

    @property
    def Z(self):
        return self._cdf_upper_bound - self._cdf_lower_bound

    mu = property(lambda self: self._mu)
    sigma = property(lambda self: self._sigma)
    lower_bound = property(lambda self: self._lower_bound)
    upper_bound = property(lambda self: self._upper_bound)
    dim = property(lambda self: self._dim)

    def __len__(self):
        return self._dim

    def sample(self, n_samples, temperature=1.0):
        sigma = self._sigma * math.sqrt(temperature)
        shape = (n_samples, *self._dim)
        if self._sampling_method == "icdf":
            u = torch.rand(shape).to(self._mu)
            r = (self._cdf_upper_bound - self._cdf_lower_bound) * u + self._cdf_lower_bound
            return self._std_icdf(r) * sigma + self._mu
        rejected = torch.ones(n_samples, device=self._mu.device, dtype=bool)
        samples = torch.empty(shape, device=self._mu.device, dtype=self._mu.dtype)
        while True:
            n_rejected = int(rejected.long().sum())
            samples[rejected] = torch.randn(n_rejected, *self._dim, device=self._mu.device,
                                            dtype=self._mu.dtype) * sigma + self._mu
            rejected = torch.any(((samples > self._upper_bound) | (samples < self._lower_bound)).view(n_samples, -1),
                                 dim=-1)
            if not torch.any(rejected):
                return samples

    def energy(self, x, temperature=1.0):
        energies = ((x - self._mu) / self._sigma) ** 2
        if self.assert_range:
            if (x < self._lower_bound).any() or (x > self._upper_bound).any():
                raise ValueError("input out of bounds")
        else:
            energies = torch.where((x < self._lower_bound) | (x > self._upper_bound),
                                   torch.full_like(energies, math.inf), energies)
        return 0.5 * energies.sum(dim=-1, keepdim=True) / temperature

    def icdf(self, x):
        return self._std_icdf(self.Z * x + self._cdf_lower_bound) * self._sigma + self._mu

    def cdf(self, x):
        return (self._std_cdf((x - self._mu) / self._sigma) - self._cdf_lower_bound) / self.Z

    def log_prob(self, x):
        z = (x - self._mu) / self._sigma
        return -0.5 * z ** 2 - 0.5 * math.log(2 * math.pi) - torch.log(self.Z * self._sigma)


class SloppyUniform(torch.nn.Module):
    """Uniform(low, high) whose argument validation tolerates ``tol`` outside the support
    (distributions.py:71-97); ``cdf / icdf / log_prob`` are torch's Uniform's."""

    def __init__(self, low, high, validate_args=None, tol=1e-5):
        super().__init__()
        self.register_buffer("low", low)
        self.register_buffer("high", high)
        self.tol = tol
        self.validate_args = validate_args

    def cdf(self, x):
        return ((x - self.low) / (self.high - self.low)).clamp(min=0, max=1)

    def icdf(self, p):
        return p * (self.high - self.low) + self.low

    def log_prob(self, x):
        inside = (self.low.le(x) & self.high.gt(x)).type_as(self.low)
        return torch.log(inside) - torch.log(self.high - self.low)

    def sample(self, sample_shape=torch.Size()):
        shape = torch.Size(sample_shape) + self.low.shape
        return self.low + torch.rand(shape, dtype=self.low.dtype, device=self.low.device) * (self.high - self.low)


# ------------------------------------------------------------------------------------------------


def _per_column(t, width):
    v = torch.as_tensor(t).detach().double().cpu().reshape(-1)
    if v.numel() == 1:
        v = v.expand(width)
    if v.numel() != width:
        raise ValueError(f"marginal parameter has {v.numel()} entries, tensor has {width} columns")
    return v.tolist()


def marginal_columns(distribution, width):
    """Column specs ``(kind, a, b, lower, upper)`` for ``engine.CdfTable`` or None if the
    distribution has no kernel (duck-typed, so the reference's own distribution objects work)."""
    d = distribution
    # (isinstance, not hasattr: the reference's SloppyUniform.__getattr__ answers None for any name)
    if all(isinstance(getattr(d, a, None), torch.Tensor) for a in ("_mu", "_logsigma", "_lower_bound", "_upper_bound")):
        mu, sg = _per_column(d._mu, width), _per_column(torch.exp(d._logsigma), width)
        lo, hi = _per_column(d._lower_bound, width), _per_column(d._upper_bound, width)
        cols = [(_lib.DIST_TRUNCNORMAL, m, s, l, h) for m, s, l, h in zip(mu, sg, lo, hi)]
        # the reference evaluates cdf / icdf / log_prob with the Phi bounds frozen at construction
        # (normal.py:148-151, possibly stale after training or a checkpoint load): hand exactly those over
        clo, chi = getattr(d, "_cdf_lower_bound", None), getattr(d, "_cdf_upper_bound", None)
        if isinstance(clo, torch.Tensor) and isinstance(chi, torch.Tensor):
            cols = [c + (a, b) for c, a, b in zip(cols, _per_column(clo, width), _per_column(chi, width))]
        return cols
    if isinstance(d, torch.distributions.Normal):
        return [(_lib.DIST_NORMAL, m, s, 0.0, 0.0)
                for m, s in zip(_per_column(d.loc, width), _per_column(d.scale, width))]
    if isinstance(d, torch.distributions.Independent):
        return marginal_columns(d.base_dist, width)
    low, high = getattr(d, "low", None), getattr(d, "high", None)
    if isinstance(low, torch.Tensor) and isinstance(high, torch.Tensor):
        return [(_lib.DIST_UNIFORM, l, h, 0.0, 0.0)
                for l, h in zip(_per_column(low, width), _per_column(high, width))]
    return None


def _dist_tensors(distribution):
    d = distribution
    names = ("_mu", "_logsigma", "_lower_bound", "_upper_bound", "_cdf_lower_bound", "_cdf_upper_bound", "loc", "scale",
             "low", "high")
    if isinstance(d, torch.distributions.Independent):
        d = d.base_dist
    return [getattr(d, n) for n in names if isinstance(getattr(d, n, None), torch.Tensor)]


def _torch_cdf_transform(distribution, x, inverse, eps):
    """cdf.py:29-46 with torch ops on the device (generic distributions, learnable marginals)."""
    if not inverse:
        y = distribution.cdf(x)
        if eps is not None:
            y = y.clamp(eps, 1.0 - eps)
        logdet = distribution.log_prob(x)
    else:
        if eps is not None:
            x = x.clamp(eps, 1.0 - eps)
        y = distribution.icdf(x)
        logdet = -distribution.log_prob(y)
    if eps is not None:
        logdet = logdet.clamp_min(-1 / eps)
    return y, logdet.sum(dim=-1, keepdim=True)


class _CdfMapFn(torch.autograd.Function):
    """Kernel forward; the backward is the closed-form derivative of an elementwise map whose
    log-det is the log of its own slope: with s = log|dy/dx| (per element),
    ``dy/dx = exp(s)`` and ``ds/dx = -/+ z / sigma * (...)`` for the Gaussian kinds, 0 for uniform."""

    @staticmethod
    def forward(ctx, owner, inverse, n, *tensors):
        outs, dlogp = engine.cdf_map(list(tensors), owner._table_for(tensors), inverse=inverse, eps=owner._eps)
        ctx.owner, ctx.inverse = owner, inverse
        ctx.save_for_backward(*tensors, *outs)
        ctx.n = n
        return (*outs, dlogp)

    @staticmethod
    def backward(ctx, *grads):
        saved = ctx.saved_tensors
        xs, ys = saved[:ctx.n], saved[ctx.n:]
        g_dl = grads[-1]
        eps = ctx.owner._eps
        gins = []
        for dist, x, y, g in zip(ctx.owner._distributions, xs, ys, grads[:-1]):
            with torch.enable_grad():
                xin = x.detach().requires_grad_(True)
                yy, ld = _torch_cdf_transform(dist, xin, ctx.inverse, eps)
                outs, gouts = [], []
                if g is not None:
                    outs.append(yy)
                    gouts.append(g)
                if g_dl is not None:
                    outs.append(ld)
                    gouts.append(g_dl)
                (gx,) = torch.autograd.grad(outs, [xin], grad_outputs=gouts, allow_unused=True)
            gins.append(gx)
        return (None, None, None, *gins)


class MultiCDFFlow(Flow):
    """``CDFTransform`` of several tensors of the flow state in one launch.  ``distributions[i]``
    maps the tensor at ``indices[i]``; forward = cdf, inverse = icdf (wrap in ``InverseFlow`` for
    the sampling direction like the builder does)."""

    _accumulates_dlogp = True

    def __init__(self, distributions, indices=None, eps=1e-7):
        super().__init__()
        self._distributions = list(distributions)
        for i, d in enumerate(self._distributions):
            if isinstance(d, torch.nn.Module):
                self.add_module(f"_dist_{i}", d)
        self._indices = list(range(len(self._distributions))) if indices is None else [int(i) for i in indices]
        self._eps = eps
        self._cache = (None, None)

    def _table_for(self, tensors):
        widths = tuple(t.shape[-1] for t in tensors)
        key = (widths, tuple((t.data_ptr(), t._version) for d in self._distributions for t in _dist_tensors(d)))
        if self._cache[0] != key:
            cols = []
            for d, w in zip(self._distributions, widths):
                c = marginal_columns(d, w)
                if c is None:
                    return None
                cols.extend(c)
            self._cache = (key, engine.CdfTable(cols))
        return self._cache[1]

    def _learnable(self):
        return any(t.requires_grad for d in self._distributions for t in _dist_tensors(d))

    def _run(self, xs, inverse, _dlogp_acc=None, **kwargs):
        xs = list(xs)
        sel = [xs[i] for i in self._indices]
        grad = torch.is_grad_enabled()
        table = self._table_for(sel) if all(t.is_cuda and t.dtype == torch.float32 for t in sel) else None
        if table is None and all(marginal_columns(d, t.shape[-1]) is not None for d, t in zip(self._distributions, sel)):
            engine.require_cuda_fp32(*sel)            # kernel-backed marginals: no CPU / fp64 fallback
        if table is None or (grad and self._learnable()):
            dlogp = _dlogp_acc
            for i, d in zip(self._indices, self._distributions):
                xs[i], dl = _torch_cdf_transform(d, xs[i], inverse, self._eps)
                dlogp = dl if dlogp is None else dlogp + dl
        elif grad and any(t.requires_grad for t in sel):
            *outs, dlogp = _CdfMapFn.apply(self, inverse, len(sel), *sel)
            for i, o in zip(self._indices, outs):
                xs[i] = o
            if _dlogp_acc is not None:
                dlogp = _dlogp_acc + dlogp
        else:
            outs, dlogp = engine.cdf_map(sel, table, inverse=inverse, eps=self._eps, dlogp_in=_dlogp_acc)
            for i, o in zip(self._indices, outs):
                xs[i] = o
        return (*xs, dlogp)

    def _forward(self, *xs, **kwargs):
        return self._run(xs, False, **kwargs)

    def _inverse(self, *xs, **kwargs):
        return self._run(xs, True, **kwargs)


class CDFTransform(MultiCDFFlow):
    """x -> cdf(x) on the distribution's support (cdf.py:13-46).  ``distribution``: anything with
    ``cdf``, ``icdf`` and ``log_prob``; ``eps`` clamps cdf values to [eps, 1-eps] and log-dets
    to >= -1/eps."""

    def __init__(self, distribution, eps=1e-7):
        super().__init__([distribution], eps=eps)

    @property
    def distribution(self):
        return self._distributions[0]


class DistributionTransferFlow(SequentialFlow):
    """source cdf, then target icdf (cdf.py:49-64)."""

    def __init__(self, source_distribution, target_distribution, eps=1e-7):
        super().__init__([CDFTransform(source_distribution, eps=eps),
                          InverseFlow(CDFTransform(target_distribution, eps=eps))])


class ConstrainGaussianFlow(Flow):
    """Normal(mu, sigma) -> the same Gaussian truncated to [lower_bound, upper_bound] (cdf.py:67-125)."""

    def __init__(self, mu, sigma=torch.tensor(1.0), lower_bound=0.0, upper_bound=math.inf, assert_range=True,
                 mu_out=None, sigma_out=None, eps=1e-7):
        super().__init__()
        source = torch.distributions.Normal(mu, sigma.to(mu))
        lower_bound, upper_bound = float(lower_bound), float(upper_bound)
        target = TruncatedNormalDistribution(
            mu=mu if mu_out is None else mu_out.to(mu), sigma=sigma if sigma_out is None else sigma_out.to(mu),
            lower_bound=lower_bound * torch.ones_like(mu), upper_bound=upper_bound * torch.ones_like(mu),
            assert_range=assert_range)
        self._trafo = DistributionTransferFlow(source, target, eps)
        self._lower_bound, self._upper_bound = lower_bound, upper_bound

    def _forward(self, x, *args, **kwargs):
        y, dlogp = self._trafo.forward(x, *args, **kwargs)
        return y.clamp(self._lower_bound, self._upper_bound), dlogp

    def _inverse(self, x, *args, **kwargs):
        return self._trafo.forward(x, *args, **kwargs, inverse=True)


# ------------------------------------------------------------------------------------------------
# the builder tail: icdf maps per field + InverseFlow(GlobalInternalCoordinateTransformation)
# ------------------------------------------------------------------------------------------------

class MappedICTail(Flow):
    """``[icdf(bonds), icdf(angles), icdf(torsions)] -> GlobalIC^-1`` as ONE kernel per direction
    (``bgx_ic_to_xyz_mapped`` / ``bgx_ic_from_xyz_mapped``).  Forward (sampling direction):
    ``(bonds, angles, torsions, x0, R) in [0,1]-space -> (xyz, dlogp)``; inverse: xyz -> the five
    tensors with bonds / angles / torsions mapped through their cdfs."""

    _accumulates_dlogp = True

    def __init__(self, coordinate_transform, marginals, eps=1e-7):
        super().__init__()
        self._ic = coordinate_transform
        self._marginals = list(marginals)            # distributions of bonds, angles, torsions
        for i, d in enumerate(self._marginals):
            if isinstance(d, torch.nn.Module):
                self.add_module(f"_marginal_{i}", d)
        self._eps = eps
        self._table = None

    def _get_table(self):
        # rebuilt when a marginal's tensors change (learnable marginals after an optimiser step, .to())
        key = tuple((t.data_ptr(), t._version) for d in self._marginals for t in _dist_tensors(d))
        if self._table is None or self._table[0] != key:
            ic = self._ic
            cols = []
            for d, w in zip(self._marginals, (ic.dim_bonds, ic.dim_angles, ic.dim_torsions)):
                cols.extend(marginal_columns(d, w))
            self._table = (key, engine.CdfTable(cols))
        return self._table[1]

    def _learnable(self):
        return any(t.requires_grad for d in self._marginals for t in _dist_tensors(d))

    def _unfused(self):
        return SequentialFlow([InverseFlow(MultiCDFFlow(self._marginals, indices=(0, 1, 2), eps=self._eps)),
                               InverseFlow(self._ic)])

    def _forward(self, bonds, angles, torsions, x0, R, _dlogp_acc=None, **kwargs):
        ins = (bonds, angles, torsions, x0, R)
        if torch.is_grad_enabled() and (any(t.requires_grad for t in ins) or self._learnable()):
            *out, dlogp = self._unfused()(*ins)
            return (*out, dlogp if _dlogp_acc is None else _dlogp_acc + dlogp)
        return engine.ic_to_xyz_mapped(self._ic._plan, self._get_table(), self._eps, *ins, dlogp_in=_dlogp_acc)

    def _inverse(self, xyz, _dlogp_acc=None, **kwargs):
        if torch.is_grad_enabled() and (xyz.requires_grad or self._learnable()):
            *out, dlogp = self._unfused()(xyz, inverse=True)
            return (*out, dlogp if _dlogp_acc is None else _dlogp_acc + dlogp)
        return engine.ic_from_xyz_mapped(self._ic._plan, self._get_table(), self._eps, xyz, dlogp_in=_dlogp_acc)


def _as_icdf_block(block):
    """(index, distribution, eps) if ``block`` is ``WrapFlow(InverseFlow(CDFTransform(d)), (i,))``."""
    if not isinstance(block, WrapFlow) or len(block._indices) != 1 or tuple(block._out_indices) != tuple(block._indices):
        return None
    inner = block._flow
    if not isinstance(inner, InverseFlow) or type(inner._delegate) is not CDFTransform:
        return None
    cdf = inner._delegate
    return int(block._indices[0]), cdf.distribution, cdf._eps


def fuse_domain_maps(flow):
    """Rewrite the tail of a builder-style ``SequentialFlow`` (generator_builder.py:408-459):

    * consecutive ``WrapFlow(InverseFlow(CDFTransform(d_i)), (i,))`` blocks become one
      ``InverseFlow(MultiCDFFlow)`` (one launch for all fields);
    * if the run is followed by ``WrapFlow(InverseFlow(GlobalIC), [b, a, t, origin, rotation], ...)``
      and maps b, a and t with kernel-backed marginals, those three maps move into the IC kernel
      (``MappedICTail``).

    Returns a new ``SequentialFlow`` sharing the original blocks' modules; flows without the
    pattern come back unchanged."""
    from .ic import GlobalInternalCoordinateTransformation
    blocks = list(flow._blocks)
    out, i = [], 0
    while i < len(blocks):
        run = []
        j = i
        while j < len(blocks):
            hit = _as_icdf_block(blocks[j])
            if hit is None or any(hit[0] == r[0] for r in run) or (run and hit[2] != run[0][2]):
                break
            run.append(hit)
            j += 1
        if not run:
            out.append(blocks[i])
            i += 1
            continue
        eps = run[0][2]
        # constants appended behind every mapped tensor (the builder's ORIGIN / ROTATION,
        # generator_builder.py:425-427) commute with the maps: hop over them
        consts = []
        while (j < len(blocks) and isinstance(blocks[j], SetConstantFlow)
               and min(blocks[j].indices) > max(r[0] for r in run)):
            consts.append(blocks[j])
            j += 1
        nxt = blocks[j] if j < len(blocks) else None
        folded = set()
        tail = None
        if (isinstance(nxt, WrapFlow) and isinstance(nxt._flow, InverseFlow)
                and isinstance(nxt._flow._delegate, GlobalInternalCoordinateTransformation)
                and len(nxt._indices) == 5):
            ic = nxt._flow._delegate
            by_index = {idx: d for idx, d, _ in run}
            b, a, t = (int(v) for v in nxt._indices[:3])
            widths = (ic.dim_bonds, ic.dim_angles, ic.dim_torsions)
            if all(k in by_index and marginal_columns(by_index[k], w) is not None
                   for k, w in zip((b, a, t), widths)):
                tail = WrapFlow(MappedICTail(ic, [by_index[b], by_index[a], by_index[t]], eps=eps),
                                nxt._indices, nxt._out_indices)
                folded = {b, a, t}
        out.extend(consts)
        rest = [(idx, d) for idx, d, _ in run if idx not in folded]
        if rest:
            out.append(InverseFlow(MultiCDFFlow([d for _, d in rest], indices=[idx for idx, _ in rest], eps=eps)))
        if tail is not None:
            out.append(tail)
            j += 1
        i = j
    return SequentialFlow(out)
