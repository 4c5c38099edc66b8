"""Run the REFERENCE's own model code in the authoring container (needs /root/reference; never used at test time).

e3nn==0.4.4, torch_scatter, torch_cluster, beartype, plotly and open3d cannot be installed here, so the reference package
does not import.  This module registers stand-ins for exactly the third-party API surface the reference's score-network
path touches, implemented on the ORACLE's restatement of those libraries (oracle/so3.py: 3j tables, harmonics, J matrices;
oracle/graph.py: radius / fps / scatter) and then imports the reference UNMODIFIED from /root/reference.  What that buys:
the reference's own module code -- graph_parser, graph_attention, gnn_block, multiscale_tensor_field, score_head,
unet_feature_extractor, score_model_base (several thousand lines the oracle restates by hand) -- becomes executable, and its
outputs pin the oracle's restatement of the REFERENCE.  The third-party arithmetic itself stays pinned only by the second
sources of tests/test_oracle_independent.py ("parity unpinned" for e3nn proper).

    from tests.golden import ref_shim; ref_shim.install(); import diffusion_edf.score_head
"""
from __future__ import annotations

import collections
import importlib.machinery
import itertools
import math
import os
import sys
import types
from typing import List, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import graph as OG      # noqa: E402
from oracle import so3              # noqa: E402

REF_ROOT = "/root/reference"


# ----------------------------------------------------------------------------------------------- o3.Irrep / o3.Irreps
class Irrep(tuple):
    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                l, p = int(s[:-1]), {"e": 1, "o": -1, "y": None}[s[-1]]
                if p is None:
                    p = (-1) ** l
            elif isinstance(l, tuple):
                l, p = l
        assert isinstance(l, int) and l >= 0 and p in (-1, 1), (l, p)
        return super().__new__(cls, (l, p))

    @property
    def l(self) -> int:      # noqa: E743
        return self[0]

    @property
    def p(self) -> int:
        return self[1]

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def is_scalar(self) -> bool:
        return self.l == 0 and self.p == 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        return [Irrep(l, p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]

    def __rmul__(self, mul: int):
        assert isinstance(mul, int)
        return Irreps([(mul, self)])

    def __add__(self, other):
        return Irreps(self) + Irreps(other)

    def __len__(self):
        raise NotImplementedError


class _MulIr(tuple):
    def __new__(cls, mul, ir=None):
        if ir is None:
            mul, ir = mul
        return super().__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self) -> int:
        return self[0]

    @property
    def ir(self) -> Irrep:
        return self[1]

    @property
    def dim(self) -> int:
        return self.mul * self.ir.dim

    def __repr__(self):
        return f"{self.mul}x{self.ir}"


class Irreps(tuple):
    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, str):
            s = irreps.replace(" ", "")
            if s:
                for tok in s.split("+"):
                    mul, ir = tok.split("x") if "x" in tok else ("1", tok)
                    out.append(_MulIr(int(mul), Irrep(ir)))
        elif irreps is None:
            pass
        else:
            for e in irreps:
                if isinstance(e, str):
                    out.append(_MulIr(1, Irrep(e)))
                elif isinstance(e, Irrep):
                    out.append(_MulIr(1, e))
                elif isinstance(e, _MulIr):
                    out.append(e)
                else:
                    mul, ir = e
                    out.append(_MulIr(mul, Irrep(ir)))
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    def slices(self) -> List[slice]:
        s, i = [], 0
        for mul_ir in self:
            s.append(slice(i, i + mul_ir.dim))
            i += mul_ir.dim
        return s

    def randn(self, *size, normalization="component", requires_grad=False, dtype=None, device=None):
        di = size.index(-1)
        shape = size[:di] + (self.dim,) + size[di + 1:]
        return torch.randn(shape, dtype=dtype, device=device, requires_grad=requires_grad)

    def __getitem__(self, i):
        x = super().__getitem__(i)
        return Irreps(x) if isinstance(i, slice) else x

    def __contains__(self, ir) -> bool:
        ir = Irrep(ir)
        return ir in (irr for _, irr in self)

    def count(self, ir) -> int:
        ir = Irrep(ir)
        return sum(mul for mul, irr in self if irr == ir)

    def index(self, _object):
        raise NotImplementedError

    def __add__(self, irreps):
        return Irreps(super().__add__(Irreps(irreps)))

    def __mul__(self, other):
        if isinstance(other, Irreps):
            raise NotImplementedError("Use o3.TensorProduct for this, see the documentation")
        return Irreps(super().__mul__(other))

    def __rmul__(self, other):
        return Irreps(super().__rmul__(other))

    def simplify(self) -> "Irreps":
        out = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def remove_zero_multiplicities(self) -> "Irreps":
        return Irreps([(mul, ir) for mul, ir in self if mul > 0])

    def sort(self):
        Ret = collections.namedtuple("sort", ["irreps", "p", "inv"])
        out = [(ir, i, mul) for i, (mul, ir) in enumerate(self)]
        out = sorted(out, key=lambda t: (t[0].l, -t[0].p, t[1]))      # e3nn orders irreps 0e, 0o, 1o, 1e, ...; only 'e' occurs here
        inv = tuple(i for _, i, _ in out)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        return Ret(Irreps([(mul, ir) for ir, _, mul in out]), tuple(p), inv)

    @property
    def dim(self) -> int:
        return sum(mul * ir.dim for mul, ir in self)

    @property
    def num_irreps(self) -> int:
        return sum(mul for mul, _ in self)

    @property
    def ls(self) -> List[int]:
        return [ir.l for mul, ir in self for _ in range(mul)]

    @property
    def lmax(self) -> int:
        return max(self.ls)

    def __repr__(self):
        return "+".join(f"{mul_ir}" for mul_ir in self)


# ----------------------------------------------------------------------------------------------- o3.TensorProduct
Instruction = collections.namedtuple("Instruction", "i_in1 i_in2 i_out connection_mode has_weight path_weight path_shape")


class TensorProduct(torch.nn.Module):
    """e3nn 0.4.4 o3.TensorProduct, the subset the reference builds: 'uvu' / 'uvw' / 'uuu' paths, irrep_normalization
    'component', path_normalization 'none' (tensor_product_rescale.py:38-42) or 'element' (the default, used by the
    ElementwiseTensorProduct of the gates); weights internal + shared, or external per sample."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions, in1_var=None, in2_var=None, out_var=None,
                 irrep_normalization=None, path_normalization=None, internal_weights=None, shared_weights=None,
                 normalization=None, compile_left_right=True, compile_right=False, _specialized_code=None, _optimize_einsums=None):
        super().__init__()
        if normalization is not None:
            irrep_normalization = normalization
        irrep_normalization = irrep_normalization or "component"
        path_normalization = path_normalization or "element"
        assert irrep_normalization == "component"
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        ins = [x if len(x) == 6 else x + (1.0,) for x in instructions]
        self.instructions = []
        for i1, i2, io, mode, has_w, pw in ins:
            m1, m2, mo = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul, self.irreps_out[io].mul
            shape = {"uvw": (m1, m2, mo), "uvu": (m1, m2), "uvv": (m1, m2), "uuw": (m1, mo), "uuu": (m1,), "uvuv": (m1, m2)}[mode]
            self.instructions.append(Instruction(i1, i2, io, mode, has_w, pw, shape))
        # path coefficients (e3nn: alpha = ir_out.dim ['component'] * path_weight / normalisation over the paths of an output)
        def num_elements(i):
            return {"uvw": self.irreps_in1[i.i_in1].mul * self.irreps_in2[i.i_in2].mul, "uvu": self.irreps_in2[i.i_in2].mul,
                    "uvv": self.irreps_in1[i.i_in1].mul, "uuw": self.irreps_in1[i.i_in1].mul, "uuu": 1, "uvuv": 1}[i.connection_mode]
        self._coef = []
        for i in self.instructions:
            alpha = self.irreps_out[i.i_out].ir.dim
            if path_normalization == "element":
                x = sum(num_elements(j) for j in self.instructions if j.i_out == i.i_out)
            elif path_normalization == "path":
                x = num_elements(i) * len([j for j in self.instructions if j.i_out == i.i_out])
            else:
                assert path_normalization == "none"
                x = 1
            if x > 0:
                alpha /= x
            alpha *= i.path_weight
            self._coef.append(math.sqrt(alpha))
        self.weight_numel = sum(math.prod(i.path_shape) for i in self.instructions if i.has_weight)
        if shared_weights is False and internal_weights is None:
            internal_weights = False
        if shared_weights is None:
            shared_weights = True
        if internal_weights is None:
            internal_weights = shared_weights and any(i.has_weight for i in self.instructions)
        assert shared_weights or not internal_weights
        self.internal_weights, self.shared_weights = internal_weights, shared_weights
        if internal_weights and self.weight_numel > 0:
            self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))
        else:
            self.register_buffer("weight", torch.Tensor())
        self.register_buffer("output_mask", torch.ones(self.irreps_out.dim))

    def _get_weights(self, weight):
        if weight is None:
            assert self.internal_weights or self.weight_numel == 0, "Weights must be provided when the TensorProduct does not have internal_weights"
            return self.weight
        if self.shared_weights:
            assert weight.shape == (self.weight_numel,), "Invalid weight shape"
        else:
            assert weight.shape[-1] == self.weight_numel and weight.ndim > 1, "Invalid weight shape"
        return weight

    def weight_views(self, weight=None, yield_instruction=False):
        weight = self._get_weights(weight)
        batchshape = weight.shape[:-1]
        off = 0
        for k, ins in enumerate(self.instructions):
            if ins.has_weight:
                n = math.prod(ins.path_shape)
                w = weight.narrow(-1, off, n).view(batchshape + ins.path_shape)
                off += n
                yield (k, ins, w) if yield_instruction else w

    def forward(self, x, y, weight=None):
        weight = self._get_weights(weight)
        lead = x.shape[:-1]
        x1 = x.reshape(-1, x.shape[-1])
        x2 = y.reshape(-1, y.shape[-1]).expand(x1.shape[0], -1) if y.shape[:-1] != lead else y.reshape(-1, y.shape[-1])
        N = x1.shape[0]
        per_sample = weight.ndim > 1
        if per_sample:
            weight = weight.reshape(-1, weight.shape[-1])
        s1, s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        outs = [x1.new_zeros(N, mul, ir.dim) for mul, ir in self.irreps_out]
        off = 0
        z = "z" if per_sample else ""
        for ins, coef in zip(self.instructions, self._coef):
            (m1, ir1), (m2, ir2), (mo, iro) = self.irreps_in1[ins.i_in1], self.irreps_in2[ins.i_in2], self.irreps_out[ins.i_out]
            w = None
            if ins.has_weight:
                n = math.prod(ins.path_shape)
                w = weight[..., off:off + n].reshape(weight.shape[:-1] + ins.path_shape)
                off += n
            if m1 == 0 or m2 == 0 or mo == 0:
                continue
            a = x1[:, s1[ins.i_in1]].reshape(N, m1, ir1.dim)
            b = x2[:, s2[ins.i_in2]].reshape(N, m2, ir2.dim)
            C = so3.wigner_3j(ir1.l, ir2.l, iro.l).to(device=x1.device, dtype=x1.dtype) * coef
            mode = ins.connection_mode
            if mode == "uvw":
                t = torch.einsum("zui,zvj,ijk->zuvk", a, b, C)
                o = torch.einsum(f"{z}uvw,zuvk->zwk", w, t)
            elif mode == "uvu":
                t = torch.einsum("zui,zvj,ijk->zuvk", a, b, C)
                o = torch.einsum(f"{z}uv,zuvk->zuk", w, t) if w is not None else t.sum(2)
            elif mode == "uuu":
                o = torch.einsum("zui,zuj,ijk->zuk", a, b, C)
                if w is not None:
                    o = o * (w[:, :, None] if per_sample else w[None, :, None])
            else:
                raise NotImplementedError(mode)
            outs[ins.i_out] = outs[ins.i_out] + o
        out = torch.cat([o.reshape(N, -1) for o in outs], dim=1) if outs else x1.new_zeros(N, 0)
        return out.reshape(lead + (self.irreps_out.dim,))


class FullyConnectedTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, irrep_normalization=None, path_normalization=None, **kwargs):
        i1, i2, io = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        instr = [(a, b, c, "uvw", True, 1.0) for a, (_, ia) in enumerate(i1) for b, (_, ib) in enumerate(i2)
                 for c, (_, ic) in enumerate(io) if ic in ia * ib]
        super().__init__(i1, i2, io, instr, irrep_normalization=irrep_normalization, path_normalization=path_normalization, **kwargs)


class ElementwiseTensorProduct(TensorProduct):
    """e3nn 0.4.4 o3.ElementwiseTensorProduct: inputs are split to equal multiplicities, one 'uuu' path per pair and output."""

    def __init__(self, irreps_in1, irreps_in2, filter_ir_out=None, irrep_normalization=None, **kwargs):
        i1, i2 = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify()
        assert i1.num_irreps == i2.num_irreps
        i1, i2 = list(i1), list(i2)
        i = 0
        while i < len(i1):
            (m1, r1), (m2, r2) = i1[i], i2[i]
            if m1 < m2:
                i2[i] = (m1, r2)
                i2.insert(i + 1, (m2 - m1, r2))
            if m2 < m1:
                i1[i] = (m2, r1)
                i1.insert(i + 1, (m1 - m2, r1))
            i += 1
        out, instr = [], []
        for i, ((mul, r1), (mul2, r2)) in enumerate(zip(i1, i2)):
            assert mul == mul2
            for ir in r1 * r2:
                if filter_ir_out is not None and ir not in [Irrep(f) for f in filter_ir_out]:
                    continue
                instr.append((i, i, len(out), "uuu", False))
                out.append((mul, ir))
        super().__init__(Irreps(i1), Irreps(i2), Irreps(out), instr, irrep_normalization=irrep_normalization, **kwargs)


# ----------------------------------------------------------------------------------------------- harmonics, misc
class SphericalHarmonics(torch.nn.Module):
    def __init__(self, irreps_out, normalize: bool, normalization: str = "integral", irreps_in=None):
        super().__init__()
        self.irreps_out = Irreps(irreps_out) if not isinstance(irreps_out, int) else Irreps([(1, (irreps_out, (-1) ** irreps_out))])
        self._ls = [ir.l for mul, ir in self.irreps_out for _ in range(mul)]
        self.normalize, self.normalization = normalize, normalization
        self.irreps_in = Irreps("1o") if irreps_in is None else Irreps(irreps_in)
        self._lmax = max(self._ls)

    def forward(self, x):
        sh = so3.spherical_harmonics(self._lmax, x, normalize=self.normalize)       # 'component'
        blocks = [sh[..., l * l:(l + 1) * (l + 1)] for l in self._ls]
        if self.normalization == "integral":
            blocks = [b / math.sqrt(4 * math.pi) for b in blocks]
        elif self.normalization == "norm":
            blocks = [b / math.sqrt(2 * l + 1) for b, l in zip(blocks, self._ls)]
        return torch.cat(blocks, dim=-1)


def spherical_harmonics(l, x, normalize, normalization="integral"):
    irr = Irreps([(1, (ll, 1)) for ll in ([l] if isinstance(l, int) else l)]) if not isinstance(l, (str, Irreps)) else Irreps(l)
    return SphericalHarmonics(irr, normalize, normalization)(x)


def _normalize2mom(f, dtype=None, device=None):
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    cst = f(z).pow(2).mean().pow(-0.5).item()

    class _N(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.f, self.cst = f, cst
            self._is_id = abs(cst - 1) < 1e-4

        def forward(self, x):
            return self.f(x) if self._is_id else self.f(x).mul(self.cst)
    return _N()


def _perm_module():
    m = types.ModuleType("e3nn.math.perm")

    def inverse(p):
        return tuple(p.index(i) for i in range(len(p)))
    m.inverse = inverse
    return m


def _direct_sum(*matrices):
    front = matrices[0].shape[:-2]
    m = sum(x.size(-2) for x in matrices)
    n = sum(x.size(-1) for x in matrices)
    out = matrices[0].new_zeros(front + (m, n))
    i = j = 0
    for x in matrices:
        out[..., i:i + x.size(-2), j:j + x.size(-1)] = x
        i += x.size(-2)
        j += x.size(-1)
    return out


def _compile_mode(mode):
    def deco(cls):
        return cls
    return deco


# ----------------------------------------------------------------------------------------------- torch_scatter / torch_cluster
def _scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    dim = dim % src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    src_m = src.movedim(dim, 0)
    idx = index.movedim(dim, 0) if index.dim() == src.dim() else index
    if idx.dim() > 1:
        idx = idx.reshape(idx.shape[0], -1)[:, 0]              # broadcast index: identical along the other dims
    if reduce in ("sum", "add"):
        res = OG.scatter_sum(src_m, idx, dim_size)
    elif reduce == "mean":
        s = OG.scatter_sum(src_m, idx, dim_size)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add_(0, idx, torch.ones(len(idx), dtype=src.dtype, device=src.device)).clamp_(min=1)
        res = s / cnt.reshape((-1,) + (1,) * (s.dim() - 1))
    else:
        raise NotImplementedError(reduce)
    return res.movedim(0, dim)


def _scatter_logsumexp(src, index, dim=-1, out=None, dim_size=None, eps=1e-12):
    dim = dim % src.dim()
    idx = index.movedim(dim, 0) if index.dim() == src.dim() else index
    if idx.dim() > 1:
        idx = idx.reshape(idx.shape[0], -1)[:, 0]
    if dim_size is None:
        dim_size = int(idx.max()) + 1
    return OG.scatter_logsumexp(src.movedim(dim, 0), idx, dim_size, eps).movedim(0, dim)


def _radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1, batch_size=None):
    return OG.radius(x, y, r, batch_x, batch_y, max_num_neighbors)


def _radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", num_workers=1, batch_size=None):
    e = OG.radius_graph(x, r, batch, loop, max_num_neighbors)
    assert flow == "source_to_target"
    return torch.stack([e[1], e[0]], dim=0)        # torch_cluster returns [col, row] = (source, target) for this flow


def _fps(src, batch=None, ratio=None, random_start=True, batch_size=None):
    return OG.fps(src, batch, float(ratio), random_start=random_start)


# ----------------------------------------------------------------------------------------------- installation
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _identity_decorator(*a, **k):
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return lambda f: f


class _Anything:
    """stand-in for modules the path never calls (plotly, open3d, ...): attribute access yields more of the same"""

    def __init__(self, name="x"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything(self._n + "." + k)

    def __call__(self, *a, **k):
        return _Anything(self._n + "()")

    def __mro_entries__(self, bases):
        return (object,)


def install():
    """Register the stand-ins and put the reference on sys.path.  Idempotent."""
    if "e3nn" in sys.modules and getattr(sys.modules["e3nn"], "_dedf_shim", False):
        return
    o3 = _mod("e3nn.o3", Irrep=Irrep, Irreps=Irreps, TensorProduct=TensorProduct,
              FullyConnectedTensorProduct=FullyConnectedTensorProduct, ElementwiseTensorProduct=ElementwiseTensorProduct,
              SphericalHarmonics=SphericalHarmonics, spherical_harmonics=spherical_harmonics, wigner_3j=so3.wigner_3j)
    wig = _mod("e3nn.o3._wigner", _Jd=[so3.J_matrix(l) if l > 0 else torch.ones(1, 1, dtype=torch.float64) for l in range(3)])
    o3._wigner = wig
    perm = _perm_module()
    sys.modules["e3nn.math.perm"] = perm
    linalg = _mod("e3nn.math._linalg", direct_sum=_direct_sum)
    emath = _mod("e3nn.math", normalize2mom=_normalize2mom, perm=perm, direct_sum=_direct_sum, _linalg=linalg)
    jit = _mod("e3nn.util.jit", compile_mode=_compile_mode, script=lambda m: m)
    argt = _mod("e3nn.util._argtools", _get_device=lambda mod: next((t.device for t in itertools.chain(mod.parameters(), mod.buffers())), torch.device("cpu")))
    util = _mod("e3nn.util", jit=jit, _argtools=argt)
    def tp_path_exists(irreps_in1, irreps_in2, ir_out):
        i1, i2, io = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify(), Irrep(ir_out)
        return any(io in a * b for _, a in i1 for _, b in i2)
    gp = _mod("e3nn.nn.models.v2106.gate_points_message_passing", tp_path_exists=tp_path_exists)
    v2106 = _mod("e3nn.nn.models.v2106", gate_points_message_passing=gp)
    models = _mod("e3nn.nn.models", v2106=v2106)
    enn = _mod("e3nn.nn", models=models)
    for m in (gp, v2106, models):
        m.__path__ = []
    e3 = _mod("e3nn", o3=o3, math=emath, util=util, nn=enn, __version__="0.4.4", _dedf_shim=True)
    e3.__path__ = []
    for m in (emath, util, o3, enn):
        m.__path__ = []
    def _named(reduce):
        def f(src, index, dim=-1, out=None, dim_size=None):
            return _scatter(src, index, dim, out, dim_size, reduce)
        return f
    _mod("torch_scatter", scatter=_scatter, scatter_sum=_named("sum"), scatter_add=_named("sum"), scatter_mean=_named("mean"),
         scatter_logsumexp=_scatter_logsumexp, scatter_softmax=None, scatter_log_softmax=None)
    _mod("torch_cluster", radius=_radius, radius_graph=_radius_graph, fps=_fps, graclus=None, knn=None)
    bt = _mod("beartype", beartype=_identity_decorator)
    bt.__path__ = []
    bt.door = _mod("beartype.door", is_bearable=lambda obj, hint: True, die_if_unbearable=lambda obj, hint, **k: None)
    bt.typing = _mod("beartype.typing")
    bt.typing.__getattr__ = lambda k: getattr(__import__("typing"), k)
    for name in ("plotly", "plotly.graph_objects", "plotly.express", "plotly.subplots", "open3d", "Pyro5", "Pyro5.api", "Pyro5.server",
                 "Pyro5.errors", "dash", "dash_vtk", "dash_daq", "dash_vtk.utils", "dash.dependencies", "dash.exceptions",
                 "jupyter_dash", "gdown", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
                 "IPython", "IPython.display", "trimesh", "pytorch3d", "wandb", "tensorboard"):
        if name not in sys.modules:
            m = _mod(name)
            m.__path__ = []
            def _ga(k, _n=name):
                if k.startswith("__"):
                    raise AttributeError(k)
                return _Anything(_n + "." + k)
            m.__getattr__ = _ga
    for p in (REF_ROOT, os.path.join(REF_ROOT, "edf_interface")):
        if p not in sys.path:
            sys.path.insert(0, p)
