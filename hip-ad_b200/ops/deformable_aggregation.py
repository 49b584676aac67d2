"""autograd boundary of the B200 deformable aggregation op.

Mirrors ``projects/mmdet3d_plugin/ops/deformable_aggregation.py:7-75`` of HiP-AD: same class
name, same ``apply(mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights)``
signature, same 5-tuple of gradients ``(g_feat, None, None, g_loc, g_w)``.  The work is done by
hand-written sm_100a kernels behind the C ABI of ``include/hipad_dfa.h`` (loaded with ctypes);
there is no other implementation to fall back to.

Differences from the reference, all behind the same interface:
  * launches go to the CURRENT torch stream (the reference uses the legacy default stream), so the
    op is correct under side streams and capturable in CUDA graphs;
  * bf16 feature maps are consumed as stored (fp32 accumulate, fp32 output) instead of being
    up-cast to a fp32 copy; every other float dtype is normalised to fp32 like the reference;
  * outputs / gradients are written in full by the kernels: no zero-fill passes, no atomics,
    bitwise reproducible run to run;
  * ``ctx.needs_input_grad`` is honoured: the dense feature gradient is skipped when the
    feature maps do not require grad (the reference always computes all three).
"""
import torch
from torch.autograd.function import Function, once_differentiable

from .. import _lib


_ACTIVE_TIMER = None


class KernelTimer:
    """Measurement aid: ``with KernelTimer() as t:`` puts CUDA events (current stream) around every launch sequence
    this package issues inside the block; ``t.summary()`` -> {"forward": ms, "backward": ms}.  Off by default."""

    def __init__(self):
        self.spans = {"forward": [], "backward": []}
        self._prev = None

    def __enter__(self):
        global _ACTIVE_TIMER
        self._prev, _ACTIVE_TIMER = _ACTIVE_TIMER, self
        return self

    def __exit__(self, *exc):
        global _ACTIVE_TIMER
        _ACTIVE_TIMER = self._prev

    def summary(self):
        torch.cuda.synchronize()
        return {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in self.spans.items()}


class _timed:
    def __init__(self, kind):
        self.kind = kind
        self.t = _ACTIVE_TIMER

    def __enter__(self):
        if self.t is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.t is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.t.spans[self.kind].append((self.a, b))


def _as_i32(t):
    cached = getattr(t, "_hipad_i32", None)
    if cached is not None and cached.device == t.device:
        return cached
    return t.contiguous().int()


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise _lib.HipadDfaError(
                "hipad_dfa: all tensors must live on a CUDA device (got %s); there is no CPU path" % t.device)


def _dims(feat, shapes, loc, weights):
    bs, num_feat, C = feat.shape
    cams, L = shapes.shape[:2]
    A, P = loc.shape[1:3]
    G = weights.shape[-1]
    if tuple(loc.shape) != (bs, A, P, cams, 2):
        raise ValueError("sampling_location must be [bs, anchors, pts, cams, 2], got %s" % (tuple(loc.shape),))
    if weights.numel() != bs * A * P * cams * L * G:
        raise ValueError("weights must be [bs, anchors, pts, cams, levels, groups], got %s" % (tuple(weights.shape),))
    return bs, cams, num_feat, C, L, A, P, G


def _norm_feat(feat):
    if feat.dtype == torch.bfloat16:
        return feat.contiguous()
    return feat.contiguous().float()


class _GradHolder:
    """One dense feature-gradient buffer shared by every aggregation call that reads the same feature tensor.
    The buffer belongs to ONE backward pass (``task``: the autograd graph-task id that created it): a pass that
    never reaches the sink node (``autograd.grad`` w.r.t. locations only, an exception, a partial graph under
    ``retain_graph``) must not leave a stale buffer for the next one."""
    __slots__ = ("buffer", "task")

    def __init__(self):
        self.buffer = None
        self.task = None


def _current_task():
    try:
        return torch._C._current_graph_task_id()
    except Exception:
        return -1


def _holder_buffer(holder):
    """The shared buffer of the CURRENT backward pass, or None (a buffer left by another pass is dropped)."""
    if holder.buffer is not None and holder.task != _current_task():
        holder.buffer = None
    return holder.buffer


def _holder_publish(holder, buffer):
    if holder.buffer is None:
        holder.task = _current_task()

        def _end_of_pass(h=holder, t=holder.task):
            if h.task == t:            # the sink did not run in this pass: nobody will read the buffer
                h.buffer = None
        try:
            torch.autograd.Variable._execution_engine.queue_callback(_end_of_pass)
        except Exception:
            pass
    holder.buffer = buffer


class _SharedFeatureGradient(Function):
    """Identity on the feature tensor.  The aggregation calls downstream add their feature gradients into ONE fp32
    buffer (hipad_dfa_group_backward with the accumulate flag) and return no gradient of their own; this node hands
    the buffer to autograd once, after all of them have run (autograd's topological order guarantees that), narrowed
    to the feature dtype in one rounding."""

    @staticmethod
    def forward(ctx, feat, holder):
        ctx.holder = holder
        ctx.feat_dtype = feat.dtype
        ctx.set_materialize_grads(False)
        return feat.view_as(feat)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        shared = _holder_buffer(ctx.holder)
        ctx.holder.buffer = None
        ctx.holder.task = None
        if shared is None:
            return grad, None
        if grad is not None:           # consumers that are not aggregation calls (e.g. the inverse-format views)
            shared = shared + grad.to(shared.dtype)
        return shared.to(ctx.feat_dtype), None


def share_feature_gradient(col_feats):
    """Returns ``col_feats`` wired so that all ``deformable_aggregation_function`` / ``deformable_aggregation_group``
    calls consuming the RETURNED tensor accumulate their feature gradients into one dense fp32 buffer: one zero fill
    per step and no per-call dense gradient (the reference materialises and autograd sums one [bs, F, C] tensor per
    call, 24 per stage-2 step).  Results are identical up to fp32 summation order, which stays deterministic
    (backward call order)."""
    if not (torch.is_tensor(col_feats) and col_feats.is_cuda and col_feats.requires_grad):
        return col_feats
    holder = _GradHolder()
    out = _SharedFeatureGradient.apply(col_feats, holder)
    out._hipad_gsink = holder
    return out


def _group_dims(feat, shapes, G):
    bs, num_feat, C = feat.shape
    cams, L = shapes.shape[:2]
    return bs, cams, num_feat, C, L, G


def _run_group_forward(lib, feat, shapes, starts, locs, ws_, out):
    """One launch for all (loc, weights) pairs.  Returns the status (ERR_UNSUPPORTED: caller goes call by call)."""
    bs, num_feat, C = feat.shape
    cams, L = shapes.shape[:2]
    G = ws_[0].shape[-1]
    table = _lib.call_table([(l.data_ptr(), w.data_ptr(), None, None, l.shape[1], l.shape[2]) for l, w in zip(locs, ws_)])
    import ctypes
    tp = ctypes.cast(table, ctypes.c_void_p)
    nbytes = lib.hipad_dfa_group_forward_workspace_bytes(tp, len(locs), bs, cams, C)
    work = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=feat.device)
    with _timed("forward"):
        return lib.hipad_dfa_group_forward(1 if feat.dtype == torch.bfloat16 else 0, out.data_ptr(), feat.data_ptr(),
                                           shapes.data_ptr(), starts.data_ptr(), tp, len(locs), bs, cams, num_feat, C, L, G,
                                           work.data_ptr(), work.numel(), torch.cuda.current_stream().cuda_stream)


def _run_group_backward(lib, feat, shapes, starts, locs, ws_, go_packed, need_feat, holder):
    """One backward chain for all calls.  Returns (status, g_feat or None, [g_loc], [g_w]); with a holder the feature
    gradient is accumulated into the pass-wide fp32 buffer and g_feat is None."""
    import ctypes
    bs, num_feat, C = feat.shape
    cams, L = shapes.shape[:2]
    G = ws_[0].shape[-1]
    bf16 = feat.dtype == torch.bfloat16
    g_locs = [torch.empty_like(l) for l in locs]
    g_ws = [torch.empty_like(w) for w in ws_]
    flags, g_feat = 0, None
    if need_feat:
        shared = _holder_buffer(holder) if holder is not None else None
        if holder is not None:
            flags |= 2                                   # shared buffers are fp32 whatever the feature type
            if shared is not None:
                g_feat, flags = shared, flags | 1
            else:
                g_feat = torch.empty(feat.shape, dtype=torch.float32, device=feat.device)
        else:
            g_feat = torch.empty_like(feat)
    table = _lib.call_table([(l.data_ptr(), w.data_ptr(), gl.data_ptr(), gw.data_ptr(), l.shape[1], l.shape[2])
                             for l, w, gl, gw in zip(locs, ws_, g_locs, g_ws)])
    tp = ctypes.cast(table, ctypes.c_void_p)
    nbytes = lib.hipad_dfa_group_backward_workspace_bytes(tp, len(locs), bs, cams, num_feat, C, L, G) if need_feat else 0
    work = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=feat.device)
    with _timed("backward"):
        rc = lib.hipad_dfa_group_backward(1 if bf16 else 0, flags, feat.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                          tp, len(locs), go_packed.data_ptr(), g_feat.data_ptr() if need_feat else None,
                                          bs, cams, num_feat, C, L, G, work.data_ptr(), work.numel(),
                                          torch.cuda.current_stream().cuda_stream)
    if rc == 0 and need_feat and holder is not None:
        _holder_publish(holder, g_feat)
        g_feat = None
    return rc, g_feat, g_locs, g_ws


class DeformableAggregationFunction(Function):
    @staticmethod
    def forward(ctx, mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights):
        lib = _lib.get()
        ctx.gsink = getattr(mc_ms_feat, "_hipad_gsink", None)
        _require_cuda(mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights)
        feat = _norm_feat(mc_ms_feat)
        shapes = _as_i32(spatial_shape)
        starts = _as_i32(scale_start_index)
        loc = sampling_location.contiguous().float()
        w = weights.contiguous().float()
        dims = _dims(feat, shapes, loc, w)
        bs, _, _, C, L, A, P, G = dims
        w = w.view(bs, A, P, dims[1], L, G)
        with torch.cuda.device(feat.device):
            out = torch.empty((bs, A, C), dtype=torch.float32, device=feat.device)
            # grouped kernel with a single call (rows of map / plan queries are cut into work units, which needs the
            # workspace this entry point takes); layouts it does not cover run on the 5-argument entry point
            rc = _run_group_forward(lib, feat, shapes, starts, [loc], [w], out)
            if rc == _lib.ERR_UNSUPPORTED:
                fn = lib.hipad_dfa_forward_bf16 if feat.dtype == torch.bfloat16 else lib.hipad_dfa_forward_f32
                rc = fn(out.data_ptr(), feat.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                        loc.data_ptr(), w.data_ptr(), *dims, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "hipad_dfa_forward")
        ctx.save_for_backward(feat, shapes, starts, loc, w)
        ctx.feat_dtype = mc_ms_feat.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        lib = _lib.get()
        feat, shapes, starts, loc, w = ctx.saved_tensors
        need_feat = ctx.needs_input_grad[0]
        go = grad_output.contiguous().float()
        holder = ctx.gsink if need_feat else None
        with torch.cuda.device(feat.device):
            rc, g_feat, g_locs, g_ws = _run_group_backward(lib, feat, shapes, starts, [loc], [w], go, need_feat, holder)
        _lib.check(rc, "hipad_dfa_backward")
        if need_feat and g_feat is not None and g_feat.dtype != ctx.feat_dtype:
            g_feat = g_feat.to(ctx.feat_dtype)
        return g_feat, None, None, g_locs[0], g_ws[0]


class DeformableAggregationGroupFunction(Function):
    """All aggregation calls of one decoder layer (they read the same feature maps) as ONE forward launch and ONE
    backward chain with ONE feature gradient.  apply(feat, shapes, starts, loc_0, w_0, loc_1, w_1, ...) returns the
    packed output [bs, sum(A_k), C]; ``deformable_aggregation_group`` splits it per call."""

    @staticmethod
    def forward(ctx, mc_ms_feat, spatial_shape, scale_start_index, *loc_w):
        lib = _lib.get()
        ctx.gsink = getattr(mc_ms_feat, "_hipad_gsink", None)
        _require_cuda(mc_ms_feat, spatial_shape, scale_start_index, *loc_w)
        feat = _norm_feat(mc_ms_feat)
        shapes = _as_i32(spatial_shape)
        starts = _as_i32(scale_start_index)
        locs = [t.contiguous().float() for t in loc_w[0::2]]
        ws_ = []
        for l, w in zip(locs, loc_w[1::2]):
            w = w.contiguous().float()
            d = _dims(feat, shapes, l, w)
            ws_.append(w.view(d[0], d[5], d[6], d[1], d[4], d[7]))
        bs, _, C = feat.shape
        a_total = sum(l.shape[1] for l in locs)
        with torch.cuda.device(feat.device):
            out = torch.empty((bs, a_total, C), dtype=torch.float32, device=feat.device)
            rc = _run_group_forward(lib, feat, shapes, starts, locs, ws_, out)
        _lib.check(rc, "hipad_dfa_group_forward")
        ctx.save_for_backward(feat, shapes, starts, *locs, *ws_)
        ctx.n = len(locs)
        ctx.feat_dtype = mc_ms_feat.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        lib = _lib.get()
        feat, shapes, starts = ctx.saved_tensors[:3]
        locs = list(ctx.saved_tensors[3:3 + ctx.n])
        ws_ = list(ctx.saved_tensors[3 + ctx.n:])
        need_feat = ctx.needs_input_grad[0]
        go = grad_output.contiguous().float()
        holder = ctx.gsink if need_feat else None
        with torch.cuda.device(feat.device):
            rc, g_feat, g_locs, g_ws = _run_group_backward(lib, feat, shapes, starts, locs, ws_, go, need_feat, holder)
        _lib.check(rc, "hipad_dfa_group_backward")
        if need_feat and g_feat is not None and g_feat.dtype != ctx.feat_dtype:
            g_feat = g_feat.to(ctx.feat_dtype)
        flat = []
        for gl, gw in zip(g_locs, g_ws):
            flat += [gl, gw]
        return (g_feat, None, None, *flat)


def group_supported(mc_ms_feat, spatial_shape, weights_list):
    """True when the grouped kernels cover this layout (4 levels, <= 8 groups, C = 128/256 f32 or 256 bf16)."""
    C = mc_ms_feat.shape[-1]
    L = spatial_shape.shape[1]
    G = weights_list[0].shape[-1]
    if L != 4 or G > 8 or C % G != 0 or any(w.shape[-1] != G for w in weights_list):
        return False
    if mc_ms_feat.dtype == torch.bfloat16:
        return C == 256 and (C // G) % 8 == 0
    return C in (128, 256) and (C // G) % 4 == 0


def deformable_aggregation_group(mc_ms_feat, spatial_shape, scale_start_index, calls):
    """calls: list of (sampling_location [bs,A_k,P_k,cams,2], weights [bs,A_k,P_k,cams,L,G]) reading the SAME feature
    maps (the det / map / plan / ego calls of one decoder layer, sparse_onedecoder.py:867-887).  Returns the list of
    outputs [bs,A_k,C], computed by one launch (and differentiated by one backward chain writing one feature
    gradient).  Layouts outside the grouped kernels' family are run call by call: same results."""
    calls = list(calls)
    if not calls:
        return []
    if len(calls) > _lib.MAX_GROUP_CALLS or not group_supported(mc_ms_feat, spatial_shape, [w for _, w in calls]):
        return [DeformableAggregationFunction.apply(mc_ms_feat, spatial_shape, scale_start_index, l, w) for l, w in calls]
    flat = []
    for l, w in calls:
        flat += [l, w]
    packed = DeformableAggregationGroupFunction.apply(mc_ms_feat, spatial_shape, scale_start_index, *flat)
    return list(packed.split([l.shape[1] for l, _ in calls], dim=1))


class AggregationWeightsFunction(Function):
    """weights_fc logits -> the op's weight tensor in ONE pass each way: group softmax over cams*L*P
    (blocks.py:196-208), the training attn-drop mask (blocks.py:209-212, drawn in the kernel instead of on the CPU) and the
    permute(0,1,4,2,3,5).contiguous() copy (blocks.py:147-158).  apply(logits, cams, L, P, G, drop_p, seed, keep_mask)
    with logits [bs, A, cams*L*P*G (any trailing shape)] -> weights [bs, A, P, cams, L, G]."""

    @staticmethod
    def forward(ctx, logits, cams, L, P, G, drop_p, seed, keep_mask):
        lib = _lib.get()
        _require_cuda(logits)
        x = logits.contiguous().float()
        bs, A = x.shape[:2]
        if x.numel() != bs * A * cams * L * P * G:
            raise ValueError("logits must hold bs*A*cams*L*P*G elements, got %s" % (tuple(x.shape),))
        mask = keep_mask.contiguous().float() if keep_mask is not None else None
        with torch.cuda.device(x.device):
            w = torch.empty((bs, A, P, cams, L, G), dtype=torch.float32, device=x.device)
            stats = torch.empty((bs, A, G, 2), dtype=torch.float32, device=x.device)
            rc = lib.hipad_dfa_weights_forward(x.data_ptr(), w.data_ptr(), stats.data_ptr(),
                                               mask.data_ptr() if mask is not None else None, int(seed), float(drop_p),
                                               bs * A, cams, L, P, G, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "hipad_dfa_weights_forward")
        ctx.save_for_backward(x, stats, mask if mask is not None else x.new_empty(0))
        ctx.meta = (cams, L, P, G, float(drop_p), int(seed), mask is not None, logits.shape, logits.dtype)
        return w

    @staticmethod
    @once_differentiable
    def backward(ctx, g_w):
        lib = _lib.get()
        x, stats, mask = ctx.saved_tensors
        cams, L, P, G, drop_p, seed, has_mask, shape, dtype = ctx.meta
        gw = g_w.contiguous().float()
        with torch.cuda.device(x.device):
            g_x = torch.empty_like(x)
            rc = lib.hipad_dfa_weights_backward(x.data_ptr(), stats.data_ptr(), gw.data_ptr(), g_x.data_ptr(),
                                                mask.data_ptr() if has_mask else None, seed, drop_p, x.shape[0] * x.shape[1],
                                                cams, L, P, G, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "hipad_dfa_weights_backward")
        return g_x.reshape(shape).to(dtype), None, None, None, None, None, None, None


def aggregation_weights(logits, cams, L, P, G, drop_p=0.0, seed=None, keep_mask=None):
    """Softmaxed (and, when drop_p > 0, attn-dropped) aggregation weights [bs, A, P, cams, L, G] from raw ``weights_fc``
    logits [bs, A, cams, L*P*G], differentiable.  ``seed`` defaults to a draw from torch's CPU generator (reproducible
    under ``torch.manual_seed``, no device sync); ``keep_mask`` [bs, A, cams, P] overrides the in-kernel draw."""
    if drop_p > 0 and seed is None and keep_mask is None:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
    return AggregationWeightsFunction.apply(logits, cams, L, P, G, drop_p, seed or 0, keep_mask)


# import-compatibility alias (ops/__init__.py:3-4 of the reference exports both names; the A800
# twin there is the same source bound to a second build, selected by GPU name string)
DeformableAggregationFunctionA800 = DeformableAggregationFunction


def fused_deformable_aggregation(feature_maps, key_points, projection_mat, image_wh, logits,
                                 return_locations=False):
    """Inference-only fused forward: projection + group softmax + aggregation in one launch.

    feature_maps : the triple from ``feature_maps_format`` ([col_feats, spatial_shape, scale_start_index])
    key_points   : [bs, A, P, 3]          (``kps_generator`` output)
    projection_mat [bs, cams, 4, 4], image_wh [bs, cams, 2] or None   (``metas``)
    logits       : [bs, A, cams, L*P*G]   raw ``weights_fc`` output, BEFORE softmax (blocks.py:196-199)
    returns      : [bs, A, C] float32  (+ sampling locations [bs, A, P, cams, 2] if requested)
    """
    lib = _lib.get()
    col_feats, spatial_shape, scale_start_index = feature_maps
    _require_cuda(col_feats, key_points, projection_mat, logits)
    feat = _norm_feat(col_feats)
    shapes = _as_i32(spatial_shape)
    starts = _as_i32(scale_start_index)
    kp = key_points.contiguous().float()
    pm = projection_mat.contiguous().float()
    wh = image_wh.contiguous().float() if image_wh is not None else None
    lg = logits.contiguous().float()
    bs, num_feat, C = feat.shape
    cams, L = shapes.shape[:2]
    A, P = kp.shape[1:3]
    G = lg.numel() // (bs * A * cams * L * P)
    if lg.numel() != bs * A * cams * L * P * G or G == 0:
        raise ValueError("logits must hold bs*A*cams*L*P*G elements, got %s" % (tuple(lg.shape),))
    dims = (bs, cams, num_feat, C, L, A, P, G)
    with torch.cuda.device(feat.device):
        out = torch.empty((bs, A, C), dtype=torch.float32, device=feat.device)
        loc = torch.empty((bs, A, P, cams, 2), dtype=torch.float32, device=feat.device) if return_locations else None
        stream = torch.cuda.current_stream().cuda_stream
        fn = lib.hipad_dfa_fused_forward_bf16 if feat.dtype == torch.bfloat16 else lib.hipad_dfa_fused_forward_f32
        with _timed("forward"):
            rc = fn(out.data_ptr(), feat.data_ptr(), shapes.data_ptr(), starts.data_ptr(), kp.data_ptr(), pm.data_ptr(),
                    wh.data_ptr() if wh is not None else None, lg.data_ptr(),
                    loc.data_ptr() if loc is not None else None, *dims, stream)
    _lib.check(rc, "hipad_dfa_fused_forward")
    return (out, loc) if return_locations else out


_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1}


def _format_call(levels, col, inverse):
    """levels: list of contiguous [bs, cams, C, H, W] CUDA tensors; col: [bs, cams*sum(HW), C].  One kernel."""
    import ctypes
    lib = _lib.get()
    bs, cams, C = levels[0].shape[:3]
    L = len(levels)
    ptrs = (ctypes.c_void_p * L)(*[t.data_ptr() for t in levels])
    hw = (ctypes.c_int32 * (2 * L))(*[int(v) for t in levels for v in t.shape[-2:]])
    with torch.cuda.device(col.device):
        rc = lib.hipad_dfa_format_features(_DTYPE_CODE[levels[0].dtype], _DTYPE_CODE[col.dtype], int(inverse),
                                           ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(hw, ctypes.c_void_p),
                                           col.data_ptr(), bs, cams, C, L, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "hipad_dfa_format_features")


class FeatureMapsFormatFunction(Function):
    """levels [bs, cams, C, H_l, W_l] (f32 or bf16, CUDA) -> col_feats [bs, cams*sum(H_l*W_l), C] in ONE transposing
    pass (``ops/__init__.py:74-103`` of the reference does cat + permute + flatten = two full copies); the backward
    is the same kernel run in the other direction.  ``out_dtype`` lets an f32 pyramid be narrowed to bf16 on the way."""

    @staticmethod
    def forward(ctx, out_dtype, *levels):
        levels = [t.contiguous() for t in levels]
        bs, cams, C = levels[0].shape[:3]
        rows = sum(int(t.shape[-2]) * int(t.shape[-1]) for t in levels)
        col = torch.empty((bs, cams * rows, C), dtype=out_dtype or levels[0].dtype, device=levels[0].device)
        _format_call(levels, col, inverse=False)
        ctx.level_shapes = [tuple(t.shape) for t in levels]
        ctx.level_dtype = levels[0].dtype
        return col

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_col):
        grad_col = grad_col.contiguous()
        if grad_col.dtype not in _DTYPE_CODE:
            grad_col = grad_col.float()
        grads = [torch.empty(shape, dtype=ctx.level_dtype, device=grad_col.device) for shape in ctx.level_shapes]
        _format_call(grads, grad_col, inverse=True)
        return (None, *grads)


def format_feature_levels(levels, out_dtype=None):
    """CUDA path of ``feature_maps_format`` for one camera group; returns col_feats."""
    return FeatureMapsFormatFunction.apply(out_dtype, *levels)


def sample_indices(spatial_shape, scale_start_index, sampling_location):
    """int32 [bs, A, P, cams, L, 6] = (valid, h_low, w_low, level_offset, corner_mask, row0) exactly as
    the kernels compute them (the bit-exact integer contract checked against the oracle)."""
    lib = _lib.get()
    _require_cuda(spatial_shape, scale_start_index, sampling_location)
    shapes = _as_i32(spatial_shape)
    starts = _as_i32(scale_start_index)
    loc = sampling_location.contiguous().float()
    bs, A, P, cams, _ = loc.shape
    L = shapes.shape[1]
    with torch.cuda.device(loc.device):
        idx = torch.empty((bs, A, P, cams, L, 6), dtype=torch.int32, device=loc.device)
        rc = lib.hipad_dfa_sample_indices(idx.data_ptr(), shapes.data_ptr(), starts.data_ptr(), loc.data_ptr(),
                                          bs, cams, L, A, P, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "hipad_dfa_sample_indices")
    return idx
