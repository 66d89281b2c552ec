"""The PWC-Lite hot path as a schedulable unit: shapes, byte accounting and one "step".

A *step* is every Correlation / flow_warp call one PWC-Lite training iteration issues, forward
and backward, on one batch of synthetic pairs (verified call counts):

  model, per direction (flows_12 and flows_21, models/pwclite.py:430-432), decoder levels 0..4
  (C = 192,128,96,64,32):
      level > 0:  x2_warp = flow_warp(x2, flow)           pwclite.py:302   (grads to x2 AND flow)
      every lvl:  out     = Correlation(x1, x2_warp)      pwclite.py:307   (grads to both)
  loss, per direction, the 4 photometric scales (losses/flow_loss.py:126-131):
      im_recons = flow_warp(im_scaled, flow, pad=border)  C = 3, gradient wrt FLOW only

  -> 10 corr fwd + 10 corr bwd + 16 warp fwd + 16 warp bwd per step (kitti_base / sintel_base).

Everything is preallocated, so a step enqueues kernels only (CUDA-graph capturable).  The convs,
flow estimators and loss arithmetic around these calls are cuDNN / elementwise PyTorch in the
reference and are out of scope here; upstream gradients are synthetic.
"""
from dataclasses import dataclass, field

import torch

from . import ops

DECODER_CHANNELS = (192, 128, 96, 64, 32)  # models/pwclite.py:183 reversed, levels 0..4
NUM_LOSS_SCALES = 4                        # w_ph_scales [1,1,1,1,0], configs/kitti_base.json
K = ops.NUM_DISPLACEMENTS


def synthetic_flow(B, h, w, sigma, gen, fine_sigma=0.05, cell=24):
    """Seeded synthetic flow with the spatial statistics of a real one: optical flow is piecewise
    smooth (PWC-Lite itself produces each level by bilinear x2 upsampling of the coarser flow plus a
    small residual, models/pwclite.py:299-301,368-371).  Coarse N(0, sigma^2) vectors on a `cell`-pixel
    lattice, bilinearly interpolated (flow gradient ~ sigma*sqrt(2)/cell = 0.18 px/px for the defaults),
    plus N(0, fine_sigma^2) per-pixel residual.  (White-noise flow,
    as used in the parity tests, scatters every warp's taps over ~20 cache sectors and measures the
    gather hardware, not the kernel.)"""
    import torch.nn.functional as F
    ch, cw = (h + cell - 1) // cell + 1, (w + cell - 1) // cell + 1
    coarse = torch.randn(B, 2, ch, cw, generator=gen) * sigma
    flow = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)
    return flow + fine_sigma * torch.randn(B, 2, h, w, generator=gen)


def conv_out(n):
    """3x3 stride-2 pad-1 conv output size (models/pwclite.py:59-63)."""
    return (n - 1) // 2 + 1


def pyramid_shapes(H, W):
    """[(C, h, w)] for decoder levels 0..4 of an H x W input (SURVEY.md 8 header)."""
    sizes = []
    h, w = H, W
    for _ in range(6):
        h, w = conv_out(h), conv_out(w)
        sizes.append((h, w))
    # feature pyramid is reversed: level 0 = coarsest (6th conv), level 4 = 2nd conv
    return [(c,) + sizes[5 - lvl] for lvl, c in enumerate(DECODER_CHANNELS)]


def loss_shapes(H, W):
    """[(3, h, w)] for the photometric scales: flows at 1, 1/2, 1/4, 1/8 resolution
    (PWCLite outputs 256x832 .. 32x104 for KITTI; flow_loss.py:128-131 resizes the images)."""
    return [(3, H >> s, W >> s) for s in range(NUM_LOSS_SCALES)]


# ---- algorithmic bytes / flops (s = 4 bytes, flow always fp32) -------------------
def corr_fwd_bytes(B, C, H, W):
    return (2 * C + K) * 4 * B * H * W


def corr_bwd_bytes(B, C, H, W):
    return (K + 4 * C) * 4 * B * H * W


def corr_fwd_flops(B, C, H, W):
    return 2 * K * C * B * H * W


def warp_fwd_bytes(B, C, H, W):
    return (2 * C * 4 + 8) * B * H * W


def warp_bwd_bytes(B, C, H, W, need_gx=True):
    return ((3 if need_gx else 2) * C * 4 + 16) * B * H * W


@dataclass
class OpRecord:
    name: str      # e.g. "corr_bwd.L4.d12"
    kind: str      # corr_fwd | corr_bwd | warp_fwd | warp_bwd | warp_bwd_flow
    fn: object
    bytes: int
    flops: int = 0
    launches: int = 1
    shape: tuple = field(default_factory=tuple)


class HotPathStep:
    """Preallocated synthetic inputs/outputs + the ordered op list of one training step."""

    def __init__(self, B, H, W, device="cuda", seed=42, flow_sigma=3.0, pad="border", flow_kind="smooth",
                 batch_directions=False):
        """batch_directions: evaluate flows_12 and flows_21 as ONE batch of 2B pairs (x1 = [f1;f2],
        x2 = [f2;f1]) like PWCLite(batch_directions=True): same math, half the launches."""
        self.pairs = B
        self.batch_directions = batch_directions
        self.dirs = ("d12",) if batch_directions else ("d12", "d21")
        if batch_directions:
            B = 2 * B
        self.B, self.H, self.W, self.pad = B, H, W, pad
        self.flow_kind = flow_kind
        self.device = torch.device(device)
        gen = torch.Generator(device="cpu").manual_seed(seed)

        def randn(*shape, scale=1.0):
            return (torch.randn(*shape, generator=gen) * scale).to(self.device)

        def flow(h, w):
            if flow_kind == "smooth":
                return synthetic_flow(B, h, w, flow_sigma, gen).to(self.device)
            return randn(B, 2, h, w, scale=flow_sigma)  # "iid": white noise, worst case for the gathers

        self.levels = pyramid_shapes(H, W)
        self.scales = loss_shapes(H, W)
        t = self.t = {}          # every tensor, by name
        self.input_names = []    # tensors a caller would have to supply (for the e2e H2D leg)
        for lvl, (C, h, w) in enumerate(self.levels):
            for n in ("f1", "f2"):
                t[f"{n}.L{lvl}"] = randn(B, C, h, w)
                self.input_names.append(f"{n}.L{lvl}")
            for d in self.dirs:
                if lvl > 0:
                    t[f"flow.L{lvl}.{d}"] = flow(h, w)
                    t[f"x2w.L{lvl}.{d}"] = torch.empty(B, C, h, w, device=self.device)
                    t[f"gfeat2.L{lvl}.{d}"] = torch.empty(B, C, h, w, device=self.device)
                    t[f"gflow.L{lvl}.{d}"] = torch.empty(B, 2, h, w, device=self.device)
                    # persistent per call site: counters zeroed once here, left zeroed by every call
                    t[f"ws.L{lvl}.{d}"] = ops.warp_bwd_workspace(t[f"f1.L{lvl}"])
                    self.input_names.append(f"flow.L{lvl}.{d}")
                t[f"corr.L{lvl}.{d}"] = torch.empty(B, K, h, w, device=self.device)
                t[f"gcorr.L{lvl}.{d}"] = randn(B, K, h, w)
                t[f"gx1.L{lvl}.{d}"] = torch.empty(B, C, h, w, device=self.device)
                t[f"gx2.L{lvl}.{d}"] = torch.empty(B, C, h, w, device=self.device)
                self.input_names.append(f"gcorr.L{lvl}.{d}")
        for s, (C, h, w) in enumerate(self.scales):
            for n in ("im1", "im2"):
                t[f"{n}.S{s}"] = torch.rand(B, C, h, w, generator=gen).to(self.device)
                self.input_names.append(f"{n}.S{s}")
            for d in self.dirs:
                t[f"lflow.S{s}.{d}"] = flow(h, w)
                t[f"recons.S{s}.{d}"] = torch.empty(B, C, h, w, device=self.device)
                t[f"grecons.S{s}.{d}"] = randn(B, C, h, w)
                t[f"glflow.S{s}.{d}"] = torch.empty(B, 2, h, w, device=self.device)
                self.input_names += [f"lflow.S{s}.{d}", f"grecons.S{s}.{d}"]
        # Every tensor a caller supplies lives in ONE device arena (256-byte aligned views), every tensor a caller
        # reads back in another: the end-to-end path then moves a step's inputs / results with a single copy each way
        # instead of ~70 small ones (per-copy overheads cost ~10 % of the PCIe rate).
        self.input_arena = self._pack(self.input_names)
        self.result_arena = self._pack(self.result_names())
        self.ops = self._build_ops()

    def _pack(self, names):
        """Move the named tensors into one flat uint8 buffer; self.t[name] becomes a view of it."""
        offs, total = [], 0
        for n in names:
            offs.append(total)
            total += (self.t[n].numel() * self.t[n].element_size() + 255) // 256 * 256
        arena = torch.empty(total, dtype=torch.uint8, device=self.device)
        for n, o in zip(names, offs):
            old = self.t[n]
            nbytes = old.numel() * old.element_size()
            view = arena[o:o + nbytes].view(old.dtype).view(old.shape)
            view.copy_(old)
            self.t[n] = view
        return arena

    # -- schedule -------------------------------------------------------------------------------
    def _build_ops(self):
        t, B, pad = self.t, self.B, self.pad
        fwd, bwd = [], []
        for d in self.dirs:
            a, b = ("f1", "f2") if d == "d12" else ("f2", "f1")
            for lvl, (C, h, w) in enumerate(self.levels):
                x1, x2 = t[f"{a}.L{lvl}"], t[f"{b}.L{lvl}"]
                shp = (B, C, h, w)
                if lvl > 0:
                    fl, x2w = t[f"flow.L{lvl}.{d}"], t[f"x2w.L{lvl}.{d}"]
                    fwd.append(OpRecord(f"warp_fwd.L{lvl}.{d}", "warp_fwd",
                                        (lambda x2=x2, fl=fl, x2w=x2w: ops.warp_fwd(x2, fl, pad, out=x2w)),
                                        warp_fwd_bytes(*shp), shape=shp))
                else:
                    x2w = x2
                out, go = t[f"corr.L{lvl}.{d}"], t[f"gcorr.L{lvl}.{d}"]
                g1, g2 = t[f"gx1.L{lvl}.{d}"], t[f"gx2.L{lvl}.{d}"]
                fwd.append(OpRecord(f"corr_fwd.L{lvl}.{d}", "corr_fwd",
                                    (lambda x1=x1, x2w=x2w, out=out: ops.corr_fwd(x1, x2w, out=out)),
                                    corr_fwd_bytes(*shp), corr_fwd_flops(*shp), shape=shp))
                # appended in reverse of execution order: the list is flipped below
                if lvl > 0:
                    gf2, gfl, ws = t[f"gfeat2.L{lvl}.{d}"], t[f"gflow.L{lvl}.{d}"], t[f"ws.L{lvl}.{d}"]
                    bwd.append(OpRecord(f"warp_bwd.L{lvl}.{d}", "warp_bwd",
                                        (lambda x2=x2, fl=fl, g2=g2, gf2=gf2, gfl=gfl, ws=ws:
                                         ops.warp_bwd(x2, fl, g2, pad, gx=gf2, gflow=gfl, workspace=ws,
                                                      workspace_clean=True)),
                                        warp_bwd_bytes(*shp), launches=2, shape=shp))
                bwd.append(OpRecord(f"corr_bwd.L{lvl}.{d}", "corr_bwd",
                                    (lambda x1=x1, x2w=x2w, go=go, g1=g1, g2=g2:
                                     ops.corr_bwd(x1, x2w, go, gx1=g1, gx2=g2)),
                                    corr_bwd_bytes(*shp), 2 * corr_fwd_flops(*shp), launches=2, shape=shp))
            im = "im2" if d == "d12" else "im1"
            for s, (C, h, w) in enumerate(self.scales):
                x, fl = t[f"{im}.S{s}"], t[f"lflow.S{s}.{d}"]
                rec, gr, gfl = t[f"recons.S{s}.{d}"], t[f"grecons.S{s}.{d}"], t[f"glflow.S{s}.{d}"]
                shp = (B, C, h, w)
                fwd.append(OpRecord(f"lwarp_fwd.S{s}.{d}", "warp_fwd",
                                    (lambda x=x, fl=fl, rec=rec: ops.warp_fwd(x, fl, pad, out=rec)),
                                    warp_fwd_bytes(*shp), shape=shp))
                bwd.append(OpRecord(f"lwarp_bwd.S{s}.{d}", "warp_bwd_flow",
                                    (lambda x=x, fl=fl, gr=gr, gfl=gfl:
                                     ops.warp_bwd(x, fl, gr, pad, need_gx=False, gflow=gfl)),
                                    warp_bwd_bytes(*shp, need_gx=False), shape=shp))
        # backward runs after the whole forward, finest level first (autograd order)
        allops = fwd + bwd[::-1]
        # per-direction chains (forward then backward of one decoder + its loss warps): the two
        # directions share only read-only inputs, so they may run on two streams (run_concurrent)
        self.chains = {d: [op for op in allops if op.name.endswith(d)] for d in self.dirs}
        return allops

    # -- execution ------------------------------------------------------------------------------
    def run(self, ops_subset=None):
        for op in (self.ops if ops_subset is None else ops_subset):
            op.fn()

    def _streams(self, n):
        pool = getattr(self, "_pool", None)
        if pool is None:
            pool = self._pool = []
        while len(pool) < n:
            pool.append(torch.cuda.Stream(device=self.device))
        return pool[:n]

    def _run_chain(self, chain, lanes):
        """Run one direction's ops in order on the current stream; maximal runs of loss-side warps
        ("lwarp_*": 4 photometric scales, independent of each other, losses/flow_loss.py:120-131) fan out
        over `lanes` (one stream per scale) and join back."""
        cur = torch.cuda.current_stream()
        marks = getattr(self, "marks", {})  # {op name: (event before, event after)} recorded on the op's stream
        i = 0
        while i < len(chain):
            if not chain[i].name.startswith("lwarp") or not lanes:
                ev = marks.get(chain[i].name)
                if ev is not None:
                    ev[0].record(cur)
                chain[i].fn()
                if ev is not None:
                    ev[1].record(cur)
                i += 1
                continue
            j = i
            while j < len(chain) and chain[j].name.startswith("lwarp"):
                j += 1
            groups = {}
            for op in chain[i:j]:
                groups.setdefault(op.name.split(".")[1], []).append(op)   # by scale "S0".."S3"
            fork = torch.cuda.Event()
            fork.record(cur)
            joins = []
            for lane, (_, grp) in zip(lanes, sorted(groups.items())):
                lane.wait_event(fork)
                with torch.cuda.stream(lane):
                    for op in grp:
                        op.fn()
                ev = torch.cuda.Event()
                ev.record(lane)
                joins.append(ev)
            for ev in joins:
                cur.wait_event(ev)
            i = j

    def run_concurrent(self, per_stream):
        """Run {"d12": [ops], "d21": [ops]} on two streams forked from / joined to the current
        stream (works under CUDA-graph capture).  PWCLite.forward calls its two decoders back to back
        (models/pwclite.py:430-432) although they are independent; this is that overlap.  Inside each
        direction the four loss scales additionally fan out over their own streams."""
        main = torch.cuda.current_stream()
        pool = self._streams(1 + 2 * NUM_LOSS_SCALES)
        side, lanes_a, lanes_b = pool[0], pool[1:1 + NUM_LOSS_SCALES], pool[1 + NUM_LOSS_SCALES:]
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            self._run_chain(per_stream.get("d21", []), lanes_b)
        self._run_chain(per_stream.get("d12", []), lanes_a)
        join = torch.cuda.Event()
        join.record(side)
        main.wait_event(join)

    def split_around(self, dom):
        """(pre, post) stream dictionaries around one op: everything that must precede `dom` in its
        own chain plus the other direction's forward goes to pre, the rest to post."""
        d = "d12" if dom.name.endswith("d12") else "d21"
        o = "d21" if d == "d12" else "d12"
        chain = self.chains[d]
        i = chain.index(dom)
        other = self.chains[o]
        nf = sum(1 for op in other if "fwd" in op.kind)
        return ({d: chain[:i], o: other[:nf]}, {d: chain[i + 1:], o: other[nf:]})

    def measure_launches(self):
        """Run the step once eagerly and read the library's own launch counter (kernels only; the
        cudaMemsetAsync of the warp-bwd tap counters is not a kernel).  Updates every op.launches."""
        for op in self.ops:
            ops.reset_launch_count()
            op.fn()
            op.launches = ops.launch_count()
        return self.launches_per_step

    @property
    def launches_per_step(self):
        return sum(op.launches for op in self.ops)

    @property
    def bytes_per_step(self):
        return sum(op.bytes for op in self.ops)

    @property
    def flops_per_step(self):
        return sum(op.flops for op in self.ops)

    def input_bytes(self):
        return sum(self.t[n].numel() * self.t[n].element_size() for n in self.input_names)

    def result_bytes(self):
        return sum(self.t[n].numel() * self.t[n].element_size() for n in self.result_names())

    def result_names(self):
        """What a caller reads back: gradients at the finest model level and loss scale."""
        top = len(self.levels) - 1
        return [f"gx1.L{top}.d12", f"gflow.L{top}.d12"] + [f"glflow.S0.{d}" for d in self.dirs]
