"""Tensor-facing launch layer over the C ABI: packs parameters (cached), builds the segment
descriptors for the flow state and launches the fused kernels on torch's current stream.

PyTorch is used here for device memory and streams only; all arithmetic of the hot path runs
in ``libbgflow_b200.so``.  Everything requires CUDA fp32 tensors — there is no CPU fallback.
"""

import ctypes as C
import functools

import numpy as np
import torch

from . import _lib

__all__ = ["PackedNet", "affine_coupling", "spline_coupling", "ic_to_xyz", "ic_from_xyz", "ZPlan",
           "require_cuda_fp32", "config", "pipeline_status", "check_pipeline_status", "CdfTable", "cdf_map",
           "ic_to_xyz_mapped", "ic_from_xyz_mapped", "RelPlan", "relic_to_xyz", "relic_from_xyz", "split_cols",
           "merge_cols"]


def require_cuda_fp32(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "bgflow_b200 runs its flows on CUDA tensors only (got a CPU tensor); "
                "there is no CPU fallback by design")
        if t.dtype != torch.float32:
            raise NotImplementedError(f"bgflow_b200 kernels are fp32; got {t.dtype}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device_guard(fn):
    """Every launch runs with the tensors' device current (the C ABI launches on the current device and
    ``_stream()`` reads that device's current stream); tensors of one call must share their device."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None

        def visit(a):
            nonlocal dev
            if isinstance(a, torch.Tensor):
                if a.is_cuda:
                    if dev is None:
                        dev = a.device
                    elif a.device != dev:
                        raise RuntimeError(f"all tensors of one bgflow_b200 call must live on the same CUDA "
                                           f"device (got {dev} and {a.device})")
            elif isinstance(a, (list, tuple)):
                for x in a:
                    visit(x)
        for a in args:
            visit(a)
        for a in kwargs.values():
            visit(a)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _dlogp_arg(dlogp_in, B, device):
    """Validate a running log-det accumulator: fp32, on ``device``, one value per sample (a broadcastable
    ``[1, 1]`` is expanded).  Returns a dense ``[B]`` tensor or None."""
    if dlogp_in is None:
        return None
    require_cuda_fp32(dlogp_in)
    if dlogp_in.device != device:
        raise RuntimeError("dlogp accumulator lives on another device than the flow state")
    d = dlogp_in.reshape(-1)
    if d.shape[0] == 1 and B != 1:
        d = d.expand(B)
    if d.shape[0] != B:
        raise ValueError("dlogp accumulator has the wrong batch size")
    return d.contiguous()


#: kernel selection: ``precision`` "bf16x3" (default: fp32 operands split exactly into bf16 terms
#: x = x1 + x2 (+ x3); three tensor-core products a1b1 + a2b1 + a1b2 with fp32 accumulation:
#: ~2^-16 relative per product, measured 6e-7 max abs error over the 8-block golden stack) or
#: "bf16x6" (three terms, six products: fp32-equivalent, twice the tensor-core time);
#: ``force_simt`` runs the shape-general fp32 SIMT kernel even where the tensor-core kernel applies.
config = {"precision": "bf16x3", "force_simt": False,
          # couplings over several / strided tensors: gather each side into one dense buffer first when
          # the batch is large, so that the two-CTAs-per-SM kernels (dense single tensors only) apply;
          # the copies cost ~5 % of a block, the one-CTA kernel ~25 %
          "gather_segments": True, "gather_min_rows": 4096,
          # spline blocks: "auto" / "pair" (the pair kernel wherever eligible: tiles in shared memory for narrow dense
          # blocks, in place in global memory otherwise), "pair_wide" (in-place tiles even where they fit shared
          # memory), "tc2" (never the pair kernel: two-CTAs-per-SM / one-CTA / SIMT kernels; A/B)
          "spline_kernel": "auto",
          # affine blocks: "auto" (two-CTAs-per-SM / one-CTA kernels for narrow blocks, the pair kernel for wide
          # ones), "pair" (the pair kernel wherever eligible), "no_pair"
          "affine_kernel": "auto",
          # conditioner GEMMs of the recompute backward (training path):
          #   "auto" (default) = "tcgen05" wherever the conditioner is a plain DenseNet whose layers have <= 128
          #       inputs or <= 128 outputs, else "fp32";
          #   "tcgen05": recompute and input gradients on bgx_linear, weight / bias gradients on bgx_gemm_tn — our own
          #       tensor-core kernels with exact two-term bf16 operand splits (fp32-class accuracy, fp32 accumulation);
          #   "fp32": torch autograd on cuBLAS fp32 (SIMT) GEMMs, the reference's semantics on a GPU;
          #   "tf32": the same with cuBLAS TF32 tensor-core GEMMs (2^-11 per product);
          #   "bf16x3": explicit backward on cuBLAS bf16 GEMMs of the split operands (fp32-class accuracy, but cuBLAS
          #       serves bf16 -> fp32-out GEMMs of these shapes with pre-Hopper kernels and the step gets SLOWER,
          #       27 ms vs 20 ms — measured, profiles/r2_train_profile_bf16x3.txt)
          "backward_gemm": "auto"}


def backward_gemm_mode():
    """``config["backward_gemm"]`` with "auto" resolved."""
    mode = config.get("backward_gemm", "auto")
    return "tcgen05" if mode == "auto" else mode

_status = {}


def pipeline_status(device):
    """Device int32 flag the tensor-core kernels raise if their internal barrier protocol
    times out (a bug, never an input property).  ``check_pipeline_status`` reads it back."""
    key = torch.device(device)
    if key.index is None:
        key = torch.device("cuda", torch.cuda.current_device())
    if key not in _status:
        _status[key] = torch.zeros(1, dtype=torch.int32, device=key)
    return _status[key]


_poll = {}


def _poll_pipeline_status(device, every=64):
    """Asynchronous check on the product path: every ``every`` tensor-core launches the flag is copied to
    pinned host memory on the launch stream; the copy issued at the PREVIOUS poll is inspected (if it has
    completed) — no host synchronisation, and a pipeline timeout surfaces as an exception at most two polls later."""
    if torch.cuda.is_current_stream_capturing():
        return                          # inside a CUDA graph capture (host.HostPipeline): the graph's owner reads the flag
    key = torch.device(device)
    st = _poll.get(key)
    if st is None:
        st = _poll[key] = {"n": 0, "host": torch.zeros(1, dtype=torch.int32).pin_memory(), "event": None}
    st["n"] += 1
    if st["n"] % every:
        return
    if st["event"] is not None:
        if not st["event"].query():
            return                      # previous copy still in flight: look again at the next poll
        if int(st["host"][0]) != 0:
            raise _lib.BgxError("tensor-core coupling kernel reported an internal pipeline timeout "
                                "(results since the last check are invalid)")
    st["host"].copy_(pipeline_status(key), non_blocking=True)
    st["event"] = torch.cuda.Event()
    st["event"].record(torch.cuda.current_stream(key))


def check_pipeline_status(device):
    """Synchronising check used by tests / smoke: raises if a kernel flagged a pipeline timeout."""
    if int(pipeline_status(device).item()) != 0:
        raise _lib.BgxError("tensor-core coupling kernel reported an internal pipeline timeout")


def _mode_flags():
    f = 0
    if config.get("force_simt"):
        f |= _lib.FLAG_FORCE_SIMT
    if config.get("precision") == "bf16x6":
        f |= _lib.FLAG_BF16X6
    elif config.get("precision") != "bf16x3":
        raise ValueError("engine.config['precision'] must be 'bf16x6' or 'bf16x3'")
    sk = config.get("spline_kernel", "auto")
    if sk == "tc2":
        f |= _lib.FLAG_NO_PAIR
    elif sk == "pair":
        f |= _lib.FLAG_PREFER_PAIR
    elif sk == "pair_wide":
        f |= _lib.FLAG_FORCE_WIDE | _lib.FLAG_PREFER_PAIR
    elif sk != "auto":
        raise ValueError("engine.config['spline_kernel'] must be 'auto', 'pair', 'pair_wide' or 'tc2'")
    return f


_ACT_CODES = {type(None): _lib.ACT_NONE, torch.nn.ReLU: _lib.ACT_RELU, torch.nn.SiLU: _lib.ACT_SILU,
              torch.nn.Tanh: _lib.ACT_TANH, torch.nn.Identity: _lib.ACT_NONE}


def activation_code(module):
    """Map the DenseNet activation module to the kernel's code, or None if unsupported."""
    return _ACT_CODES.get(type(module))


class PackedNet:
    """Device copy of one conditioner MLP in kernel layout, cached on parameter versions."""

    def __init__(self):
        self._key = None
        self._buf = None
        self._event = None
        self._pack_stream = None
        self._seen = set()
        self._sources = None
        self.packed = _lib.bgx_packed_mlp()

    @_device_guard
    def refresh(self, weights, biases, act_code, periodic=None, spline=None):
        """weights[i]: [out,in] fp32 CUDA (nn.Linear.weight), biases[i]: [out].
        periodic: None or (indices, left, right, raw_width); spline: None or (d_t, n_bins, circ mask)."""
        key = tuple((w.data_ptr(), w._version, b.data_ptr(), b._version) for w, b in zip(weights, biases))
        key = (key, act_code, repr(periodic), repr(spline), weights[0].device)
        if key == self._key:
            cur = torch.cuda.current_stream(weights[0].device)
            if cur != self._pack_stream and cur not in self._seen:
                cur.wait_event(self._event)       # first use on another stream than the packing one
                self._seen.add(cur)
            return self.packed
        lib = _lib.load()
        require_cuda_fp32(*weights, *biases)
        n = len(weights)
        if n > _lib.BGX_MAX_LAYERS:
            raise NotImplementedError(f"at most {_lib.BGX_MAX_LAYERS} Linear layers per conditioner")
        ws = [w.detach().contiguous() for w in weights]
        bs = [b.detach().contiguous() for b in biases]
        src = _lib.bgx_mlp()
        src.n_layers = n
        src.act = act_code
        src.dims[0] = ws[0].shape[1]
        for i, (w, b) in enumerate(zip(ws, bs)):
            if w.shape[1] != src.dims[i] or b.shape[0] != w.shape[0]:
                raise ValueError("inconsistent DenseNet layer shapes")
            src.dims[i + 1] = w.shape[0]
            src.W[i] = w.data_ptr()
            src.b[i] = b.data_ptr()
        keep = []
        if periodic is not None:
            idx, left, right, raw_width = periodic
            arr = (C.c_int32 * len(idx))(*[int(i) for i in idx])
            keep.append(arr)
            src.n_periodic = len(idx)
            src.periodic_idx = C.cast(arr, C.POINTER(C.c_int32))
            src.periodic_left = float(left)
            src.periodic_right = float(right)
            src.raw_width = int(raw_width)
        else:
            src.raw_width = src.dims[0]
        lay = None
        if spline is not None:
            d_t, n_bins, circ = spline
            lay = _lib.bgx_spline_layout()
            lay.d_t = int(d_t)
            lay.n_bins = int(n_bins)
            if circ is not None:
                carr = (C.c_uint8 * d_t)(*[1 if c else 0 for c in circ])
                keep.append(carr)
                lay.is_circular = C.cast(carr, C.POINTER(C.c_uint8))
        out = _lib.bgx_packed_mlp()
        lay_p = C.byref(lay) if lay is not None else None
        rc = lib.bgx_pack_mlp(C.byref(src), lay_p, None, 0, C.byref(out), _stream())
        if rc == -1 and spline is not None:
            # mirrors the RuntimeError of spline.py:112-121 (split/reshape of a wrong-width output)
            raise RuntimeError(
                f"params_net output width {src.dims[n]} does not match "
                f"3 * n_bins * {spline[0]} + n_noncircular")
        _lib.check(rc, "bgx_pack_mlp(size)")
        buf = torch.empty(int(out.total_floats), dtype=torch.float32, device=ws[0].device)
        rc = lib.bgx_pack_mlp(C.byref(src), lay_p, C.c_void_p(buf.data_ptr()), buf.numel(), C.byref(out),
                              _stream())
        _lib.check(rc, "bgx_pack_mlp")
        if self._buf is not None:
            for st in self._seen:                 # launches on other streams may still read the old layout
                self._buf.record_stream(st)
        self._pack_stream = torch.cuda.current_stream(ws[0].device)
        self._event = torch.cuda.Event()
        self._event.record(self._pack_stream)
        self._seen = set()
        self._buf, self.packed, self._key = buf, out, key
        # the cache key is (storage address, version): hold on to the sources so that a freed temporary's address
        # cannot come back with other contents and alias a stale pack
        self._sources = (list(weights), list(biases))
        return out


def _as_rows(t):
    """View a flow-state tensor as [B, w] with unit inner stride (copy only if it must)."""
    w = t.shape[-1]
    if t.dim() != 2:
        t = t.reshape(-1, w)
    if t.stride(-1) != 1 or (t.shape[0] > 1 and t.stride(0) < w):
        t = t.contiguous()
    return t


def _fill_io(cond, tr, dlogp_in):
    if len(cond) > _lib.BGX_MAX_SEGS or len(tr) > _lib.BGX_MAX_SEGS or len(tr) < 1:
        raise NotImplementedError(f"at most {_lib.BGX_MAX_SEGS} tensors per side of a coupling")
    require_cuda_fp32(*cond, *tr)
    batch_shape = tr[0].shape[:-1]
    cond2 = [_as_rows(t) for t in cond]
    tr2 = [_as_rows(t) for t in tr]
    B = tr2[0].shape[0]
    for t in cond2 + tr2:
        if t.shape[0] != B:
            raise ValueError("all tensors of a coupling must share their batch shape")
    widths = [t.shape[1] for t in tr2]
    if (config.get("gather_segments") and not config.get("force_simt") and config.get("precision") == "bf16x3"
            and B >= config.get("gather_min_rows", 4096) and B % 4 == 0):
        def dense(t):
            return t.stride(0) == t.shape[1] and t.stride(1) == 1 and t.data_ptr() % 16 == 0
        if len(cond2) > 1 or (cond2 and not dense(cond2[0])):
            cond2 = [torch.cat(cond2, dim=1) if len(cond2) > 1 else cond2[0].contiguous()]
        if len(tr2) > 1 or not dense(tr2[0]):
            tr2 = [torch.cat(tr2, dim=1) if len(tr2) > 1 else tr2[0].contiguous()]
    out = torch.empty(B, sum(widths), dtype=torch.float32, device=tr2[0].device)
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=tr2[0].device)
    io = _lib.bgx_coupling_io()
    io.batch = B
    io.n_cond = len(cond2)
    for i, t in enumerate(cond2):
        io.cond[i].ptr, io.cond[i].width, io.cond[i].stride = t.data_ptr(), t.shape[1], t.stride(0) if B > 1 else t.shape[1]
    io.n_tr = len(tr2)
    col = 0
    outs = []
    for i, w in enumerate(widths):
        o = out[:, col:col + w]
        outs.append(o.reshape(*batch_shape, w) if len(batch_shape) != 1 else o)
        if len(tr2) == len(widths):
            t = tr2[i]
            io.tr_in[i].ptr, io.tr_in[i].width, io.tr_in[i].stride = t.data_ptr(), w, t.stride(0) if B > 1 else w
            io.tr_out[i].ptr, io.tr_out[i].width, io.tr_out[i].stride = o.data_ptr(), w, out.stride(0)
        col += w
    if len(tr2) != len(widths):          # gathered: one dense segment in, the whole output buffer out
        io.tr_in[0].ptr, io.tr_in[0].width, io.tr_in[0].stride = tr2[0].data_ptr(), col, col
        io.tr_out[0].ptr, io.tr_out[0].width, io.tr_out[0].stride = out.data_ptr(), col, col
    keep = [cond2, tr2]
    d = _dlogp_arg(dlogp_in, B, tr2[0].device)
    if d is not None:
        keep.append(d)
        io.dlogp_in = d.data_ptr()
    io.dlogp_out = dlogp.data_ptr()
    return io, outs, dlogp.reshape(*batch_shape, 1), keep


@_device_guard
def affine_coupling(cond, tr, shift, scale, log_alpha, inverse=False, preserve_volume=False,
                    is_circular=False, dlogp_in=None, flags=0):
    """cond / tr: lists of tensors (concatenated along the last dim by the kernel);
    shift / scale: ``bgx_packed_mlp`` or None.  Returns (list of outputs, dlogp [.., 1])."""
    lib = _lib.load()
    io, outs, dlogp, keep = _fill_io(cond, tr, dlogp_in)
    if io.batch == 0:
        return outs, dlogp
    # (the affine entry point has no per-call status argument: point the library at THIS device's flag)
    lib.bgx_set_status_buffer(C.c_void_p(pipeline_status(tr[0].device).data_ptr()))
    f = flags | (_mode_flags() & ~(_lib.FLAG_NO_PAIR | _lib.FLAG_FORCE_WIDE | _lib.FLAG_PREFER_PAIR)) \
        | (_lib.FLAG_INVERSE if inverse else 0) | (_lib.FLAG_PRESERVE_VOLUME if preserve_volume else 0) \
        | (_lib.FLAG_CIRCULAR if is_circular else 0)
    ak = config.get("affine_kernel", "auto")
    if ak == "pair":
        f |= _lib.FLAG_PREFER_PAIR
    elif ak == "no_pair":
        f |= _lib.FLAG_NO_PAIR
    elif ak != "auto":
        raise ValueError("engine.config['affine_kernel'] must be 'auto', 'pair' or 'no_pair'")
    rc = lib.bgx_affine_coupling(C.byref(io), C.byref(shift) if shift is not None else None,
                                 C.byref(scale) if scale is not None else None, float(log_alpha), f, _stream())
    _lib.check(rc, "bgx_affine_coupling")
    _poll_pipeline_status(tr[0].device)
    return outs, dlogp


@_device_guard
def spline_coupling(cond, tr, net, n_bins, inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0,
                    min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3, identity_init=True,
                    oob_counter=None, dlogp_in=None, flags=0):
    lib = _lib.load()
    io, outs, dlogp, keep = _fill_io(cond, tr, dlogp_in)
    if io.batch == 0:
        return outs, dlogp
    cfg = _lib.bgx_spline_cfg()
    cfg.n_bins = n_bins
    cfg.left, cfg.right, cfg.bottom, cfg.top = float(left), float(right), float(bottom), float(top)
    cfg.min_bin_width, cfg.min_bin_height, cfg.min_derivative = min_bin_width, min_bin_height, min_derivative
    cfg.identity_init = 1 if identity_init else 0
    cfg.oob_counter = oob_counter.data_ptr() if oob_counter is not None else None
    cfg.status = pipeline_status(tr[0].device).data_ptr()
    f = flags | _mode_flags() | (_lib.FLAG_INVERSE if inverse else 0)
    rc = lib.bgx_spline_coupling(C.byref(io), C.byref(net), C.byref(cfg), f, _stream())
    _lib.check(rc, "bgx_spline_coupling")
    _poll_pipeline_status(tr[0].device)
    return outs, dlogp


@_device_guard
def spline_backward(params, y, g_out, g_dlogp, end_slope_col, n_bins, inverse=False, left=0.0, right=1.0,
                    bottom=0.0, top=1.0, min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3,
                    identity_init=True):
    """dP, dy of the spline transform (``bgx_spline_backward``).  ``params`` ``[B, 3*K*D + n_nc]``
    is the conditioner output, ``end_slope_col`` an int32 device tensor with D entries."""
    lib = _lib.load()
    require_cuda_fp32(params, y, g_out)
    params, y, g_out = params.contiguous(), y.contiguous(), g_out.contiguous()
    B, d_t = y.shape
    d_params = torch.empty_like(params)
    d_y = torch.empty_like(y)
    if B == 0:
        return d_params, d_y
    gd = g_dlogp.reshape(-1).contiguous() if g_dlogp is not None else None
    cfg = _lib.bgx_spline_cfg()
    cfg.n_bins = n_bins
    cfg.left, cfg.right, cfg.bottom, cfg.top = float(left), float(right), float(bottom), float(top)
    cfg.min_bin_width, cfg.min_bin_height, cfg.min_derivative = min_bin_width, min_bin_height, min_derivative
    cfg.identity_init = 1 if identity_init else 0
    rc = lib.bgx_spline_backward(B, d_t, params.data_ptr(), params.stride(0), y.data_ptr(), g_out.data_ptr(),
                                 gd.data_ptr() if gd is not None else None, end_slope_col.data_ptr(),
                                 C.byref(cfg), _lib.FLAG_INVERSE if inverse else 0, d_params.data_ptr(),
                                 d_y.data_ptr(), _stream())
    _lib.check(rc, "bgx_spline_backward")
    return d_params, d_y


@_device_guard
def split_bf16(x):
    """Exact two-term bf16 split ``x = hi + lo`` of an fp32 CUDA tensor (``bgx_split_bf16``).  Returns two
    contiguous bf16 tensors of ``x``'s shape."""
    lib = _lib.load()
    require_cuda_fp32(x)
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if x.numel():
        rc = lib.bgx_split_bf16(C.c_void_p(x.data_ptr()), x.numel(), C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()),
                                _stream())
        _lib.check(rc, "bgx_split_bf16")
    return hi, lo


class LinearTC:
    """One linear layer ``y = x W^T + b`` on the tensor cores (``bgx_linear``; training path).  ``W`` ``[N, K]`` and
    ``b`` ``[N]`` (or None) are packed once per (storage, version)."""

    def __init__(self):
        self._net = PackedNet()
        self._zero = None

    @_device_guard
    def __call__(self, x, weight, bias=None):
        lib = _lib.load()
        require_cuda_fp32(x, weight)
        n, k = weight.shape
        if x.dim() != 2 or x.shape[1] != k:
            raise ValueError("LinearTC: x must be [B, K] for a weight [N, K]")
        if bias is None:
            if self._zero is None or self._zero.numel() != n or self._zero.device != x.device:
                self._zero = torch.zeros(n, dtype=torch.float32, device=x.device)
            bias = self._zero
        packed = self._net.refresh([weight], [bias], _lib.ACT_NONE)
        x = x.contiguous()
        y = torch.empty(x.shape[0], n, dtype=torch.float32, device=x.device)
        if x.shape[0]:
            rc = lib.bgx_linear(x.shape[0], C.c_void_p(x.data_ptr()), C.byref(packed), C.c_void_p(y.data_ptr()),
                                C.c_void_p(pipeline_status(x.device).data_ptr()), _stream())
            _lib.check(rc, "bgx_linear")
        return y

    @staticmethod
    def supports(n, k):
        return n <= 128 or k <= 128


def _pad4(n):
    return (n + 3) // 4 * 4


_train_workspaces = {}      # (device, stream, layout) -> (flat, views, bgx_train_buffers): scratch of the block backwards


class TrainNet:
    """Training-path state of one conditioner (a plain ``DenseNet``): its layers packed for the tensor-core linear
    kernel — forward and transposed (``bgx_train_pack``), cached on parameter versions — plus the two host calls that
    run its whole recompute (``bgx_mlp_forward_train``) and its whole backward (``bgx_mlp_backward``)."""

    def __init__(self):
        self._key = None
        self._buf = None
        self._event = None
        self._pack_stream = None
        self._seen = set()
        self._sources = None
        self.packed = _lib.bgx_train_mlp()
        self.dims = None

    @_device_guard
    def refresh(self, weights, biases, act_codes):
        key = tuple((w.data_ptr(), w._version, b.data_ptr(), b._version) for w, b in zip(weights, biases))
        key = (key, tuple(act_codes), weights[0].device)
        if key == self._key:
            cur = torch.cuda.current_stream(weights[0].device)
            if cur != self._pack_stream and cur not in self._seen:
                cur.wait_event(self._event)
                self._seen.add(cur)
            return self.packed
        lib = _lib.load()
        require_cuda_fp32(*weights, *biases)
        n = len(weights)
        if n > _lib.BGX_MAX_LAYERS:
            raise NotImplementedError(f"at most {_lib.BGX_MAX_LAYERS} Linear layers per conditioner")
        ws = [w.detach().contiguous() for w in weights]
        bs = [b.detach().contiguous() for b in biases]
        src = _lib.bgx_mlp()
        src.n_layers = n
        src.dims[0] = ws[0].shape[1]
        for i, (w, b) in enumerate(zip(ws, bs)):
            if w.shape[1] != src.dims[i] or b.shape[0] != w.shape[0]:
                raise ValueError("inconsistent DenseNet layer shapes")
            src.dims[i + 1] = w.shape[0]
            src.W[i] = w.data_ptr()
            src.b[i] = b.data_ptr()
        src.raw_width = src.dims[0]
        acts = (C.c_int32 * max(n - 1, 1))(*[int(a) for a in act_codes[:n - 1]])
        out = _lib.bgx_train_mlp()
        rc = lib.bgx_train_pack(C.byref(src), acts, None, 0, C.byref(out), None)
        _lib.check(rc, "bgx_train_pack(size)")
        buf = torch.empty(int(out.total_floats), dtype=torch.float32, device=ws[0].device)
        rc = lib.bgx_train_pack(C.byref(src), acts, C.c_void_p(buf.data_ptr()), buf.numel(), C.byref(out), _stream())
        _lib.check(rc, "bgx_train_pack")
        if self._buf is not None:
            for st in self._seen:
                self._buf.record_stream(st)
        self._pack_stream = torch.cuda.current_stream(ws[0].device)
        self._event = torch.cuda.Event()
        self._event.record(self._pack_stream)
        self._seen = set()
        self._buf, self.packed, self._key = buf, out, key
        self._sources = (ws, bs, list(weights), list(biases))
        self.dims = [int(src.dims[i]) for i in range(n + 1)]
        return out

    def _buffers(self, B, device, extra=(), scratch=False):
        """One allocation for a recompute + backward: z_i, then h_i and g_i of the hidden layers, the weight-gradient
        partials, then ``extra`` sizes.  Returns (flat tensor, views, bgx_train_buffers).

        ``scratch``: nothing of the allocation outlives the call that asks for it (the block-backward entry points: every
        result is copied out or lives in ``extra`` buffers of its own), so one workspace per (device, stream, layout)
        is shared by all blocks of that shape — they run one after the other on the stream — instead of a fresh
        450 MB allocation and a dozen views per block and step."""
        lib = _lib.load()
        dims = self.dims
        L = len(dims) - 1
        widths = [_pad4(d) for d in dims[1:]]
        part = int(lib.bgx_mlp_train_part_floats(B, C.byref(self.packed))) if B else 0
        sizes = [B * w for w in widths] + [B * w for w in widths[:-1]] * 2 + [part] + [int(e) for e in extra]
        sizes = [(sz + 3) // 4 * 4 for sz in sizes]          # keep every view 16-byte aligned
        key = None
        if scratch and not torch.cuda.is_current_stream_capturing():     # a graph must own what it captured
            dev = torch.device(device)
            key = (dev, torch.cuda.current_stream(dev).cuda_stream, tuple(sizes), L)
            hit = _train_workspaces.get(key)
            if hit is not None:
                return hit
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=device)
        views, off = [], 0
        for sz in sizes:
            views.append(flat[off:off + sz])
            off += sz
        bufs = _lib.bgx_train_buffers()
        for i in range(L):
            bufs.z[i] = views[i].data_ptr()
            if i + 1 < L:
                bufs.h[i] = views[L + i].data_ptr()
                bufs.g[i] = views[2 * L - 1 + i].data_ptr()
        bufs.part = views[2 * L - 1 + L - 1].data_ptr()
        if key is not None:
            if len(_train_workspaces) >= 8:
                _train_workspaces.pop(next(iter(_train_workspaces)))
            _train_workspaces[key] = (flat, views, bufs)
        return flat, views, bufs

    def _grad_buffers(self, device, zero=False):
        dims = self.dims
        L = len(dims) - 1
        sizes = []
        for i in range(L):
            sizes += [dims[i + 1] * dims[i], dims[i + 1]]
        sizes = [(sz + 3) // 4 * 4 for sz in sizes]
        flat = (torch.zeros if zero else torch.empty)(sum(sizes), dtype=torch.float32, device=device)
        grads, off = [], 0
        d_w = (C.c_void_p * L)()
        d_b = (C.c_void_p * L)()
        for i in range(L):
            w = flat[off:off + dims[i + 1] * dims[i]].view(dims[i + 1], dims[i])
            off += sizes[2 * i]
            b = flat[off:off + dims[i + 1]]
            off += sizes[2 * i + 1]
            d_w[i], d_b[i] = w.data_ptr(), b.data_ptr()
            grads += [w, b]
        return grads, d_w, d_b

    @_device_guard
    def spline_block_backward(self, x, y, g_out, g_dlogp, end_slope_col, n_bins, inverse=False, left=0.0, right=1.0,
                              bottom=0.0, top=1.0, min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3,
                              identity_init=True):
        """The whole backward of a spline coupling block whose conditioner this is, in ONE host call
        (``bgx_spline_coupling_backward``): ``x`` ``[B, dims[0]]`` the conditioner input, ``y`` ``[B, d_t]`` the block's
        transformed input, ``g_out`` / ``g_dlogp`` the upstream gradients.  Returns ``(d_x, d_y, [dW_0, db_0, ...])``."""
        lib = _lib.load()
        require_cuda_fp32(x, y, g_out)
        x, y, g_out = x.contiguous(), y.contiguous(), g_out.contiguous()
        B, d_t = y.shape
        dims = self.dims
        grads, d_w, d_b = self._grad_buffers(x.device, zero=(B == 0))
        wl = _pad4(dims[-1])
        flat, views, bufs = self._buffers(B, x.device, extra=(B * wl,), scratch=True)
        d_p = views[-1]
        res = torch.empty(B * d_t + B * dims[0], dtype=torch.float32, device=x.device)      # handed to autograd: fresh
        d_y, d_x = res[:B * d_t].view(B, d_t), res[B * d_t:].view(B, dims[0])
        if B == 0:
            return d_x, d_y, grads
        gd = g_dlogp.reshape(-1).contiguous() if g_dlogp is not None else None
        cfg = _lib.bgx_spline_cfg()
        cfg.n_bins = n_bins
        cfg.left, cfg.right, cfg.bottom, cfg.top = float(left), float(right), float(bottom), float(top)
        cfg.min_bin_width, cfg.min_bin_height, cfg.min_derivative = min_bin_width, min_bin_height, min_derivative
        cfg.identity_init = 1 if identity_init else 0
        rc = lib.bgx_spline_coupling_backward(
            B, C.byref(self.packed), C.c_void_p(x.data_ptr()), d_t, C.c_void_p(y.data_ptr()), C.c_void_p(g_out.data_ptr()),
            C.c_void_p(gd.data_ptr()) if gd is not None else None, C.c_void_p(end_slope_col.data_ptr()), C.byref(cfg),
            _lib.FLAG_INVERSE if inverse else 0, C.byref(bufs), C.c_void_p(d_p.data_ptr()), C.c_void_p(d_x.data_ptr()),
            C.c_void_p(d_y.data_ptr()), d_w, d_b, C.c_void_p(pipeline_status(x.device).data_ptr()), _stream())
        _lib.check(rc, "bgx_spline_coupling_backward")
        return d_x, d_y, grads

    @_device_guard
    def forward(self, x):
        """Recompute on ``x`` ``[B, dims[0]]``.  Returns the state ``backward`` needs; ``state["out_padded"]`` is the net's
        output ``[B, pad4(dims[-1])]`` (zero pad columns), ``state["out"]`` its first ``dims[-1]`` columns."""
        lib = _lib.load()
        require_cuda_fp32(x)
        x = x.contiguous()
        B, dims = x.shape[0], self.dims
        L = len(dims) - 1
        widths = [_pad4(d) for d in dims[1:]]
        flat, views, bufs = self._buffers(B, x.device)
        if B:
            rc = lib.bgx_mlp_forward_train(B, C.byref(self.packed), C.c_void_p(x.data_ptr()), C.byref(bufs),
                                           C.c_void_p(pipeline_status(x.device).data_ptr()), _stream())
            _lib.check(rc, "bgx_mlp_forward_train")
        out_padded = views[L - 1].view(B, widths[-1])
        return {"x": x, "flat": flat, "bufs": bufs, "out_padded": out_padded, "out": out_padded[:, :dims[-1]],
                "n_out": dims[-1], "packed": self.packed, "dims": dims}

    @_device_guard
    def backward(self, state, d_out, need_dx=True):
        """``d_out`` ``[B, dims[-1]]`` or ``[B, pad4(dims[-1])]`` with zero pad columns.  Returns ``(d_x or None,
        [dW_0, db_0, dW_1, db_1, ...])``."""
        lib = _lib.load()
        x, dims = state["x"], state["dims"]
        B, L = x.shape[0], len(dims) - 1
        wl = _pad4(dims[-1])
        if d_out.shape[1] != wl:
            d_out = torch.nn.functional.pad(d_out, (0, wl - d_out.shape[1]))
        d_out = d_out.contiguous()
        grads, d_w, d_b = self._grad_buffers(x.device, zero=(B == 0))
        d_x = torch.empty(B, dims[0], dtype=torch.float32, device=x.device) if need_dx else None
        if B:
            rc = lib.bgx_mlp_backward(B, C.byref(state["packed"]), C.c_void_p(x.data_ptr()), C.byref(state["bufs"]),
                                      C.c_void_p(d_out.data_ptr()), C.c_void_p(d_x.data_ptr()) if need_dx else None,
                                      d_w, d_b, C.c_void_p(pipeline_status(x.device).data_ptr()), _stream())
            _lib.check(rc, "bgx_mlp_backward")
        return d_x, grads


@_device_guard
def affine_block_backward(tn_shift, tn_scale, log_alpha, x, y, g_out, g_dlogp, inverse=False):
    """The whole backward of a plain RealNVP coupling block in ONE host call (``bgx_affine_coupling_backward``):
    ``tn_shift`` / ``tn_scale`` the refreshed ``TrainNet`` of its two conditioners, ``log_alpha`` the block's device
    scalar, ``x`` ``[B, K0]`` the conditioner input, ``y`` ``[B, d_t]`` the block's transformed input, ``g_out`` /
    ``g_dlogp`` the upstream gradients.  Returns ``(d_x, d_y, shift grads, scale grads, d_log_alpha[1])``."""
    lib = _lib.load()
    require_cuda_fp32(x, y, g_out, log_alpha)
    x, y, g_out = x.contiguous(), y.contiguous(), g_out.contiguous()
    B, d_t = y.shape
    k0 = x.shape[1]
    g_mu, dw_mu, db_mu = tn_shift._grad_buffers(x.device, zero=(B == 0))
    g_s, dw_s, db_s = tn_scale._grad_buffers(x.device, zero=(B == 0))
    d_la = torch.zeros(1, dtype=torch.float32, device=x.device)
    scratch = int(lib.bgx_affine_backward_scratch_floats(B, d_t, k0)) if B else 0
    # the two nets need workspaces of their own (both recomputes are alive at once): the scratch key carries the extras
    flat_mu, views_mu, bufs_mu = tn_shift._buffers(B, x.device, extra=(scratch,), scratch=True)
    flat_s, _, bufs_s = tn_scale._buffers(B, x.device, extra=(4,), scratch=True)
    res = torch.empty(B * d_t + B * k0, dtype=torch.float32, device=x.device)                # handed to autograd: fresh
    d_y, d_x = res[:B * d_t].view(B, d_t), res[B * d_t:].view(B, k0)
    if B == 0:
        return d_x, d_y, g_mu, g_s, d_la
    gd = g_dlogp.reshape(-1).contiguous() if g_dlogp is not None else None
    rc = lib.bgx_affine_coupling_backward(
        B, C.byref(tn_shift.packed), C.byref(tn_scale.packed), C.c_void_p(log_alpha.data_ptr()), C.c_void_p(x.data_ptr()),
        d_t, C.c_void_p(y.data_ptr()), C.c_void_p(g_out.data_ptr()), C.c_void_p(gd.data_ptr()) if gd is not None else None,
        _lib.FLAG_INVERSE if inverse else 0, C.byref(bufs_mu), C.byref(bufs_s), C.c_void_p(views_mu[-1].data_ptr()),
        C.c_void_p(d_x.data_ptr()), C.c_void_p(d_y.data_ptr()), dw_mu, db_mu, dw_s, db_s, C.c_void_p(d_la.data_ptr()),
        C.c_void_p(pipeline_status(x.device).data_ptr()), _stream())
    _lib.check(rc, "bgx_affine_coupling_backward")
    return d_x, d_y, g_mu, g_s, d_la


@_device_guard
def gemm_tn(g, h, n=None):
    """Weight gradient of a linear layer on the tensor cores (``bgx_gemm_tn``; training path):
    ``dW = g[:, :n]^T h`` ``[n, K]`` and ``db = g[:, :n].sum(0)`` for ``g`` ``[B, >= n]`` (row stride free),
    ``h`` ``[B, K]``, ``K <= 128``.  The kernel writes one partial per batch slice; they are added here in a fixed
    order."""
    lib = _lib.load()
    require_cuda_fp32(g, h)
    if g.dim() != 2 or h.dim() != 2 or g.shape[0] != h.shape[0] or g.stride(1) != 1 or h.stride(1) != 1:
        raise ValueError("gemm_tn: g [B, N] and h [B, K] with unit column stride")
    n = g.shape[1] if n is None else int(n)
    B, k = h.shape
    if k > 128 or n > g.shape[1]:
        raise ValueError("gemm_tn: K <= 128 and n <= g.shape[1]")
    if B == 0:
        return g.new_zeros(n, k), g.new_zeros(n)
    slices = lib.bgx_gemm_tn_slices(B, n)
    if slices <= 0:
        raise RuntimeError("bgx_gemm_tn_slices failed (no CUDA device?)")
    rows = (n + 127) // 128 * 128
    part_w = torch.empty(slices, rows, 128, dtype=torch.float32, device=g.device)
    part_b = torch.empty(slices, rows, dtype=torch.float32, device=g.device)
    rc = lib.bgx_gemm_tn(B, C.c_void_p(g.data_ptr()), g.stride(0), n, C.c_void_p(h.data_ptr()), h.stride(0), k, slices,
                         C.c_void_p(part_w.data_ptr()), C.c_void_p(part_b.data_ptr()),
                         C.c_void_p(pipeline_status(g.device).data_ptr()), _stream())
    _lib.check(rc, "bgx_gemm_tn")
    return part_w.sum(dim=0)[:n, :k], part_b.sum(dim=0)[:n]


class ZPlan:
    """Host staging of a global z-matrix (what ic.py:25-97 does with numpy) + its device copy."""

    def __init__(self, z_matrix, normalize_angles=True, eps=1e-7):
        z = np.asarray(z_matrix, dtype=np.int64)
        if z.ndim != 2 or z.shape[1] != 4:
            raise ValueError("z_matrix must have shape (n_atoms, 4)")
        undefined = (z == -1).sum(axis=1)
        seeds = []
        for want in (3, 2, 1):
            rows = np.flatnonzero(undefined == want)
            if len(rows) != 1:
                raise ValueError("a global z-matrix needs exactly one row with 3, 2 and 1 undefined references")
            seeds.append(int(z[rows[0], 0]))
        rel = z[undefined == 0]
        if len(rel) != len(z) - 3:
            raise ValueError("z-matrix rows must have 0, 1, 2 or 3 undefined (-1) references")
        n_atoms = len(z)
        if sorted(z[:, 0].tolist()) != list(range(n_atoms)):
            raise ValueError("z-matrix must place every atom exactly once")
        # topological order: a row is placeable once its three reference atoms are placed
        placed = np.zeros(n_atoms, dtype=bool)
        placed[seeds] = True
        todo = list(range(len(rel)))
        order = []
        while todo:
            ready = [r for r in todo if placed[rel[r, 1:]].all()]
            if not ready:
                raise ValueError(
                    "Z-matrix decomposition failed. The following atoms were not reachable from "
                    f"the fixed atoms: \n{rel[todo, 0]}")
            order.extend(ready)
            placed[rel[ready, 0]] = True
            ready_set = set(ready)
            todo = [r for r in todo if r not in ready_set]
        self.z_matrix = z
        self.seeds = seeds
        self.rel = rel
        self.order = order
        self.n_atoms = n_atoms
        self.normalize_angles = bool(normalize_angles)
        self.eps = float(eps)
        self._dev = {}

    def device_plan(self, device):
        key = str(device)
        if key not in self._dev:
            rel = torch.as_tensor(self.rel.astype(np.int32)).contiguous().to(device)
            order = torch.as_tensor(np.asarray(self.order, dtype=np.int32)).to(device)
            # IC column -> slot of the atom that owns it (bonds | angles | torsions column order)
            s0, s1, s2 = self.seeds
            atoms = [int(a) for a in self.rel[:, 0]]
            slots = ([3 * s1, 3 * s2] + [3 * a for a in atoms] + [3 * s2 + 1] + [3 * a + 1 for a in atoms]
                     + [3 * a + 2 for a in atoms])
            slot_of_col = torch.as_tensor(np.asarray(slots, dtype=np.int32)).to(device)
            plan = _lib.bgx_zplan()
            plan.n_atoms = self.n_atoms
            for i in range(3):
                plan.seeds[i] = self.seeds[i]
            plan.n_rel = len(self.rel)
            plan.rel = rel.data_ptr()
            plan.order = order.data_ptr()
            plan.slot_of_col = slot_of_col.data_ptr()
            plan.normalize_angles = 1 if self.normalize_angles else 0
            plan.eps = self.eps
            self._dev[key] = (plan, rel, order, slot_of_col)
        return self._dev[key][0]


@_device_guard
def ic_to_xyz(plan, bonds, angles, torsions, x0, R, dlogp_in=None):
    lib = _lib.load()
    require_cuda_fp32(bonds, angles, torsions, x0, R)
    n = plan.n_atoms
    B = bonds.shape[0]
    if bonds.shape != (B, n - 1) or angles.shape != (B, n - 2) or torsions.shape != (B, n - 3):
        raise ValueError("bonds/angles/torsions must be [B, N-1], [B, N-2], [B, N-3]")
    bonds, angles, torsions = bonds.contiguous(), angles.contiguous(), torsions.contiguous()
    x0f = x0.reshape(-1, 3).contiguous()
    Rf = R.reshape(-1, 3).contiguous()
    if x0f.shape[0] not in (1, B) or Rf.shape[0] not in (1, B):
        raise ValueError("x0 must be [B,1,3] (or [1,3]) and R [B,3]")
    xyz = torch.empty(B, 3 * n, dtype=torch.float32, device=bonds.device)
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=bonds.device)
    din = _dlogp_arg(dlogp_in, B, bonds.device)
    if B == 0:
        return xyz, dlogp
    rc = lib.bgx_ic_to_xyz(C.byref(plan.device_plan(bonds.device)), B, bonds.data_ptr(), angles.data_ptr(),
                           torsions.data_ptr(), x0f.data_ptr(), 3 if x0f.shape[0] == B else 0,
                           Rf.data_ptr(), 3 if Rf.shape[0] == B else 0, xyz.data_ptr(),
                           din.data_ptr() if din is not None else None, dlogp.data_ptr(), _stream())
    _lib.check(rc, "bgx_ic_to_xyz")
    return xyz, dlogp


@_device_guard
def ic_from_xyz(plan, xyz, dlogp_in=None):
    lib = _lib.load()
    require_cuda_fp32(xyz)
    n = plan.n_atoms
    B = xyz.shape[0]
    x = xyz.reshape(B, -1).contiguous()
    if x.shape[1] != 3 * n:
        raise ValueError(f"xyz must have {3 * n} coordinates per sample")
    dev = x.device
    bonds = torch.empty(B, n - 1, dtype=torch.float32, device=dev)
    angles = torch.empty(B, n - 2, dtype=torch.float32, device=dev)
    torsions = torch.empty(B, n - 3, dtype=torch.float32, device=dev)
    x0 = torch.empty(B, 1, 3, dtype=torch.float32, device=dev)
    R = torch.empty(B, 3, dtype=torch.float32, device=dev)
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=dev)
    din = _dlogp_arg(dlogp_in, B, dev)
    if B == 0:
        return bonds, angles, torsions, x0, R, dlogp
    rc = lib.bgx_ic_from_xyz(C.byref(plan.device_plan(dev)), B, x.data_ptr(), bonds.data_ptr(), angles.data_ptr(),
                             torsions.data_ptr(), x0.data_ptr(), R.data_ptr(),
                             din.data_ptr() if din is not None else None, dlogp.data_ptr(), _stream())
    _lib.check(rc, "bgx_ic_from_xyz")
    return bonds, angles, torsions, x0, R, dlogp


# ------------------------------------------------------------------------------------------------
# IC-domain CDF maps (bgx_cdf_map) and the fused builder tail (bgx_ic_*_mapped)
# ------------------------------------------------------------------------------------------------

def _clamp_args(eps):
    """CDFTransform's eps (cdf.py:22-27) as the three kernel scalars."""
    if eps is None:
        return 0.0, 1.0, float("-inf")
    return float(np.float32(eps)), float(np.float32(1.0 - eps)), -1.0 / eps


class CdfTable:
    """Device table of ``bgx_cdf_col`` entries, one per tensor column.

    ``columns``: sequence of ``(kind, a, b, lower, upper[, cdf_lower, cdf_upper])`` with kind in
    ``_lib.DIST_NONE / DIST_NORMAL / DIST_TRUNCNORMAL / DIST_UNIFORM`` (see ``bgx_cdf_col_init``)."""

    def __init__(self, columns):
        lib = _lib.load()
        self.n = len(columns)
        arr = (_lib.bgx_cdf_col * max(self.n, 1))()
        for i, col in enumerate(columns):
            kind, a, b, lower, upper = col[:5]
            rc = lib.bgx_cdf_col_init(int(kind), float(a), float(b), float(lower), float(upper), C.byref(arr[i]))
            if rc == 0 and len(col) == 7:      # frozen Phi(alpha), Phi(beta) of a truncated normal
                rc = lib.bgx_cdf_col_set_truncation(C.byref(arr[i]), float(col[5]), float(col[6]))
            if rc != 0:
                raise ValueError(f"invalid marginal for column {i}: kind={kind}, a={a}, b={b}, "
                                 f"lower={lower}, upper={upper}")
        self._host = torch.from_numpy(np.frombuffer(arr, dtype=np.uint8).copy())
        self._dev = {}

    def device(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = self._host.to(device)
        return self._dev[key]


@_device_guard
def cdf_map(tensors, table, inverse=False, eps=1e-7, dlogp_in=None):
    """Map a list of ``[B, w_i]`` tensors through their per-column CDFs (``inverse``: icdfs) in one
    launch.  Returns (list of mapped tensors, dlogp ``[..., 1]``)."""
    lib = _lib.load()
    if not 1 <= len(tensors) <= _lib.BGX_MAX_SEGS:
        raise NotImplementedError(f"1..{_lib.BGX_MAX_SEGS} tensors per CDF-map launch")
    require_cuda_fp32(*tensors)
    batch_shape = tensors[0].shape[:-1]
    rows = [_as_rows(t) for t in tensors]
    B = rows[0].shape[0]
    if any(t.shape[0] != B for t in rows):
        raise ValueError("all tensors of a CDF map must share their batch shape")
    if sum(t.shape[1] for t in rows) != table.n:
        raise ValueError(f"CDF table has {table.n} columns, tensors have {sum(t.shape[1] for t in rows)}")
    outs = [torch.empty(B, t.shape[1], dtype=torch.float32, device=t.device) for t in rows]
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=rows[0].device)
    segs_in = (_lib.bgx_seg * _lib.BGX_MAX_SEGS)()
    segs_out = (_lib.bgx_seg * _lib.BGX_MAX_SEGS)()
    for i, (t, o) in enumerate(zip(rows, outs)):
        w = t.shape[1]
        segs_in[i].ptr, segs_in[i].width, segs_in[i].stride = t.data_ptr(), w, t.stride(0) if B > 1 else w
        segs_out[i].ptr, segs_out[i].width, segs_out[i].stride = o.data_ptr(), w, w
    din = _dlogp_arg(dlogp_in, B, rows[0].device)
    if B > 0:
        lo, hi, ldmin = _clamp_args(eps)
        cols = table.device(rows[0].device)
        rc = lib.bgx_cdf_map(B, len(rows), segs_in, segs_out, C.c_void_p(cols.data_ptr()), lo, hi, ldmin,
                             _lib.FLAG_INVERSE if inverse else 0,
                             din.data_ptr() if din is not None else None, dlogp.data_ptr(), _stream())
        _lib.check(rc, "bgx_cdf_map")
    outs = [o.reshape(*batch_shape, o.shape[1]) if len(batch_shape) != 1 else o for o in outs]
    return outs, dlogp.reshape(*batch_shape, 1)


@_device_guard
def ic_to_xyz_mapped(plan, table, eps, bonds, angles, torsions, x0, R, dlogp_in=None):
    """icdf maps of (bonds | angles | torsions) + IC -> Cartesian in one kernel."""
    lib = _lib.load()
    require_cuda_fp32(bonds, angles, torsions, x0, R)
    n = plan.n_atoms
    B = bonds.shape[0]
    if bonds.shape != (B, n - 1) or angles.shape != (B, n - 2) or torsions.shape != (B, n - 3):
        raise ValueError("bonds/angles/torsions must be [B, N-1], [B, N-2], [B, N-3]")
    if table.n != 3 * n - 6:
        raise ValueError(f"CDF table must have {3 * n - 6} columns (bonds | angles | torsions)")
    bonds, angles, torsions = bonds.contiguous(), angles.contiguous(), torsions.contiguous()
    x0f = x0.reshape(-1, 3).contiguous()
    Rf = R.reshape(-1, 3).contiguous()
    if x0f.shape[0] not in (1, B) or Rf.shape[0] not in (1, B):
        raise ValueError("x0 must be [B,1,3] (or [1,3]) and R [B,3]")
    xyz = torch.empty(B, 3 * n, dtype=torch.float32, device=bonds.device)
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=bonds.device)
    din = _dlogp_arg(dlogp_in, B, bonds.device)
    if B == 0:
        return xyz, dlogp
    lo, hi, ldmin = _clamp_args(eps)
    cols = table.device(bonds.device)
    rc = lib.bgx_ic_to_xyz_mapped(C.byref(plan.device_plan(bonds.device)), C.c_void_p(cols.data_ptr()), lo, hi, ldmin,
                                  B, bonds.data_ptr(), angles.data_ptr(), torsions.data_ptr(), x0f.data_ptr(),
                                  3 if x0f.shape[0] == B else 0, Rf.data_ptr(), 3 if Rf.shape[0] == B else 0,
                                  xyz.data_ptr(), din.data_ptr() if din is not None else None, dlogp.data_ptr(),
                                  _stream())
    _lib.check(rc, "bgx_ic_to_xyz_mapped")
    return xyz, dlogp


@_device_guard
def ic_from_xyz_mapped(plan, table, eps, xyz, dlogp_in=None):
    """Cartesian -> IC + cdf maps of every IC column in one kernel."""
    lib = _lib.load()
    require_cuda_fp32(xyz)
    n = plan.n_atoms
    B = xyz.shape[0]
    x = xyz.reshape(B, -1).contiguous()
    if x.shape[1] != 3 * n:
        raise ValueError(f"xyz must have {3 * n} coordinates per sample")
    if table.n != 3 * n - 6:
        raise ValueError(f"CDF table must have {3 * n - 6} columns (bonds | angles | torsions)")
    dev = x.device
    bonds = torch.empty(B, n - 1, dtype=torch.float32, device=dev)
    angles = torch.empty(B, n - 2, dtype=torch.float32, device=dev)
    torsions = torch.empty(B, n - 3, dtype=torch.float32, device=dev)
    x0 = torch.empty(B, 1, 3, dtype=torch.float32, device=dev)
    R = torch.empty(B, 3, dtype=torch.float32, device=dev)
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=dev)
    din = _dlogp_arg(dlogp_in, B, dev)
    if B == 0:
        return bonds, angles, torsions, x0, R, dlogp
    lo, hi, ldmin = _clamp_args(eps)
    cols = table.device(dev)
    rc = lib.bgx_ic_from_xyz_mapped(C.byref(plan.device_plan(dev)), C.c_void_p(cols.data_ptr()), lo, hi, ldmin, B,
                                    x.data_ptr(), bonds.data_ptr(), angles.data_ptr(), torsions.data_ptr(),
                                    x0.data_ptr(), R.data_ptr(), din.data_ptr() if din is not None else None,
                                    dlogp.data_ptr(), _stream())
    _lib.check(rc, "bgx_ic_from_xyz_mapped")
    return bonds, angles, torsions, x0, R, dlogp


# ------------------------------------------------------------------------------------------------
# relative / mixed internal coordinates (bgx_relic_*)
# ------------------------------------------------------------------------------------------------

class RelPlan:
    """Host staging of a relative z-matrix (ic.py:25-91 with fixed atoms) + optional static
    whitening of the fixed block (pca.py:10-34), and their device copies."""

    def __init__(self, z_matrix, fixed_atoms, normalize_angles=True, eps=1e-7, whitening=None):
        z = np.asarray(z_matrix, dtype=np.int64)
        fixed = np.asarray(fixed_atoms, dtype=np.int64).reshape(-1)
        if z.ndim != 2 or z.shape[1] != 4:
            raise ValueError("z_matrix must have shape (n_conditioned, 4)")
        if (z < 0).any():
            raise ValueError("a relative z-matrix has no undefined (-1) references")
        n_atoms = len(z) + len(fixed)
        if sorted(z[:, 0].tolist() + fixed.tolist()) != list(range(n_atoms)):
            raise ValueError("z-matrix rows and fixed atoms must cover every atom exactly once")
        placed = np.zeros(n_atoms, dtype=bool)
        placed[fixed] = True
        todo = list(range(len(z)))
        order = []
        while todo:
            ready = [r for r in todo if placed[z[r, 1:]].all()]
            if not ready:
                raise ValueError(
                    "Z-matrix decomposition failed. The following atoms were not reachable from "
                    f"the fixed atoms: \n{z[todo, 0]}")
            order.extend(ready)
            placed[z[ready, 0]] = True
            ready_set = set(ready)
            todo = [r for r in todo if r not in ready_set]
        self.rel, self.fixed, self.order, self.n_atoms = z, fixed, order, n_atoms
        self.seeds = [int(a) for a in fixed]          # "already placed" atoms for the torch backward definition
        self.normalize_angles = bool(normalize_angles)
        self.eps = float(eps)
        self.whitening = whitening                    # None or dict(mean, whiten, blacken, jacobian_xz, keepdims)
        self._dev = {}

    @property
    def keepdims(self):
        return 0 if self.whitening is None else int(self.whitening["keepdims"])

    @property
    def fixed_width(self):
        return self.keepdims if self.whitening is not None else 3 * len(self.fixed)

    def device_plan(self, device):
        key = str(device)
        if key not in self._dev:
            as_i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(device)
            rel, fixed, order = as_i32(self.rel), as_i32(self.fixed), as_i32(self.order)
            plan = _lib.bgx_relplan()
            plan.n_atoms, plan.n_fixed, plan.n_rel = self.n_atoms, len(self.fixed), len(self.rel)
            plan.fixed, plan.rel, plan.order = fixed.data_ptr(), rel.data_ptr(), order.data_ptr()
            plan.normalize_angles = 1 if self.normalize_angles else 0
            plan.eps = self.eps
            keep = [rel, fixed, order]
            if self.whitening is not None:
                w = self.whitening
                as_f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(device)
                mean, blacken, whiten = as_f32(w["mean"]), as_f32(w["blacken"]), as_f32(w["whiten"])
                plan.keepdims = int(w["keepdims"])
                plan.mean, plan.blacken, plan.whiten = mean.data_ptr(), blacken.data_ptr(), whiten.data_ptr()
                plan.log_det_whiten = float(w["jacobian_xz"])
                keep += [mean, blacken, whiten]
            self._dev[key] = (plan, keep)
        return self._dev[key][0]


@_device_guard
def relic_to_xyz(plan, bonds, angles, torsions, fixed, dlogp_in=None):
    lib = _lib.load()
    require_cuda_fp32(bonds, angles, torsions, fixed)
    n_rel = len(plan.rel)
    B = bonds.shape[0]
    fixed2 = fixed.reshape(B, -1).contiguous()
    if bonds.shape != (B, n_rel) or angles.shape != (B, n_rel) or torsions.shape != (B, n_rel):
        raise ValueError("bonds/angles/torsions must be [B, n_conditioned]")
    if fixed2.shape[1] != plan.fixed_width:
        raise ValueError(f"the fixed block must have {plan.fixed_width} columns")
    bonds, angles, torsions = bonds.contiguous(), angles.contiguous(), torsions.contiguous()
    xyz = torch.empty(B, 3 * plan.n_atoms, dtype=torch.float32, device=bonds.device)
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=bonds.device)
    din = _dlogp_arg(dlogp_in, B, bonds.device)
    if B == 0:
        return xyz, dlogp
    rc = lib.bgx_relic_to_xyz(C.byref(plan.device_plan(bonds.device)), B, bonds.data_ptr(), angles.data_ptr(),
                              torsions.data_ptr(), fixed2.data_ptr(), xyz.data_ptr(),
                              din.data_ptr() if din is not None else None, dlogp.data_ptr(), _stream())
    _lib.check(rc, "bgx_relic_to_xyz")
    return xyz, dlogp


@_device_guard
def relic_from_xyz(plan, xyz, dlogp_in=None):
    lib = _lib.load()
    require_cuda_fp32(xyz)
    B = xyz.shape[0]
    x = xyz.reshape(B, -1).contiguous()
    if x.shape[1] != 3 * plan.n_atoms:
        raise ValueError(f"xyz must have {3 * plan.n_atoms} coordinates per sample")
    dev, n_rel = x.device, len(plan.rel)
    bonds = torch.empty(B, n_rel, dtype=torch.float32, device=dev)
    angles = torch.empty(B, n_rel, dtype=torch.float32, device=dev)
    torsions = torch.empty(B, n_rel, dtype=torch.float32, device=dev)
    fixed = torch.empty(B, plan.fixed_width, dtype=torch.float32, device=dev)
    dlogp = torch.empty(B, 1, dtype=torch.float32, device=dev)
    din = _dlogp_arg(dlogp_in, B, dev)
    if B == 0:
        return bonds, angles, torsions, fixed, dlogp
    rc = lib.bgx_relic_from_xyz(C.byref(plan.device_plan(dev)), B, x.data_ptr(), bonds.data_ptr(), angles.data_ptr(),
                                torsions.data_ptr(), fixed.data_ptr(), din.data_ptr() if din is not None else None,
                                dlogp.data_ptr(), _stream())
    _lib.check(rc, "bgx_relic_from_xyz")
    return bonds, angles, torsions, fixed, dlogp


# ------------------------------------------------------------------------------------------------
# SplitFlow / MergeFlow by sizes along the last dim (bgx_split_merge)
# ------------------------------------------------------------------------------------------------

def _seg(t, B):
    s = _lib.bgx_seg()
    s.ptr, s.width, s.stride = t.data_ptr(), t.shape[1], t.stride(0) if B > 1 else t.shape[1]
    return s


@_device_guard
def split_cols(x, sizes):
    """``x [.., W]`` -> dense tensors of the given last-dim sizes, one launch."""
    lib = _lib.load()
    require_cuda_fp32(x)
    if not 1 <= len(sizes) <= _lib.BGX_MAX_SEGS:
        raise NotImplementedError(f"1..{_lib.BGX_MAX_SEGS} parts per split")
    lead = x.shape[:-1]
    x2 = _as_rows(x)
    B = x2.shape[0]
    parts = [torch.empty(B, int(w), dtype=torch.float32, device=x.device) for w in sizes]
    if B > 0:
        segs = (_lib.bgx_seg * _lib.BGX_MAX_SEGS)(*[_seg(p, B) for p in parts])
        whole = _seg(x2, B)
        _lib.check(lib.bgx_split_merge(B, C.byref(whole), len(parts), segs, 0, _stream()), "bgx_split_merge")
    return [p.reshape(*lead, p.shape[1]) if len(lead) != 1 else p for p in parts]


@_device_guard
def merge_cols(parts):
    """Concatenate tensors along the last dim, one launch."""
    lib = _lib.load()
    require_cuda_fp32(*parts)
    if not 1 <= len(parts) <= _lib.BGX_MAX_SEGS:
        raise NotImplementedError(f"1..{_lib.BGX_MAX_SEGS} parts per merge")
    lead = parts[0].shape[:-1]
    rows = [_as_rows(p) for p in parts]
    B = rows[0].shape[0]
    if any(r.shape[0] != B for r in rows):
        raise ValueError("all tensors of a merge must share their batch shape")
    out = torch.empty(B, sum(r.shape[1] for r in rows), dtype=torch.float32, device=rows[0].device)
    if B > 0:
        segs = (_lib.bgx_seg * _lib.BGX_MAX_SEGS)(*[_seg(r, B) for r in rows])
        whole = _seg(out, B)
        _lib.check(lib.bgx_split_merge(B, C.byref(whole), len(rows), segs, 1, _stream()), "bgx_split_merge")
    return out.reshape(*lead, out.shape[1]) if len(lead) != 1 else out
