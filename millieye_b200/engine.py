"""Execution plan of the Darknet forward on the B200 kernels.

A plan is built once per (batch, image size): BN is folded and weights are packed into the GEMM
kernel's layout, every block gets an NHWC fp16 view (buffer, channel offset, pitch) - route/concat
and shortcut are resolved at plan time into channel slices and a fused residual add, so the only
kernels left are conv GEMMs, the 3-channel first conv, max-pool, upsample and the YOLO decode -
and the launch sequence is captured in a CUDA graph that later forwards replay.
"""
import gc
import os

import torch

from . import ops
from ._lib import ME_ACT_LEAKY, ME_ACT_LINEAR, MeError


class View:
    """NHWC fp16/fp32 activation view: `c` channels starting at channel `off` of `buf` (pitch = buf.shape[-1])."""

    def __init__(self, buf, off, c, h, w, real_c=None):
        self.buf, self.off, self.c, self.h, self.w = buf, off, c, h, w
        self.real_c = real_c if real_c is not None else c
        self.pitch = buf.shape[-1]

    @property
    def t(self):
        """Tensor whose data_ptr is the first element of the view."""
        return self.buf.view(-1)[self.off:]

    def at(self, b0):
        """Same, starting at frame b0 of the batch (sub-batch execution)."""
        return self.buf[b0:].view(-1)[self.off:]


def describe_blocks(module_defs):
    """Static per-block description (same rules as create_modules, reference yolov3/models.py:12-79)."""
    hyper = module_defs[0]
    chans = [int(hyper["channels"])]
    blocks = []
    for i, d in enumerate(module_defs[1:]):
        kind = d["type"]
        b = {"type": kind}
        if kind == "convolutional":
            b.update(bn=int(d["batch_normalize"]), filters=int(d["filters"]), size=int(d["size"]),
                     stride=int(d["stride"]), leaky=(d["activation"] == "leaky"), cin=chans[-1])
            out_c = b["filters"]
        elif kind == "maxpool":
            b.update(size=int(d["size"]), stride=int(d["stride"]))
            out_c = chans[-1]
        elif kind == "upsample":
            b.update(stride=int(d["stride"]))
            out_c = chans[-1]
        elif kind == "route":
            b.update(layers=[int(v) if int(v) >= 0 else i + int(v) for v in d["layers"].split(",")])
            out_c = sum(chans[1:][j] for j in b["layers"])
        elif kind == "shortcut":
            b.update(src=i + int(d["from"]))
            out_c = chans[1:][b["src"]]
        elif kind == "yolo":
            mask = [int(v) for v in d["mask"].split(",")]
            flat = [int(v) for v in d["anchors"].split(",")]
            pairs = list(zip(flat[0::2], flat[1::2]))
            b.update(anchors=[pairs[j] for j in mask], classes=int(d["classes"]))
            out_c = chans[-1]
        else:
            raise MeError(f"cfg block type '{kind}' is not part of the supported Darknet subset")
        b["out_c"] = out_c
        blocks.append(b)
        chans.append(out_c)
    return hyper, blocks


def capture_graph(fn):
    """Captures fn() into a CUDA graph with the cyclic garbage collector paused.  A collection that runs in the
    middle of a capture can finalise an older plan's CUDAGraph (plans sit in reference cycles with their modules);
    cudaGraphExecDestroy is not permitted while a stream is capturing and invalidates the capture in progress."""
    g = torch.cuda.CUDAGraph()
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            fn()
    finally:
        if was_enabled:
            gc.enable()
    return g


class DarknetPlan:
    """Buffers, packed weights and the op list for one (n, size) shape on one device."""

    def __init__(self, blocks, tensors, n, size, device, feature_tap, in_channels=3, splits=None):
        self.blocks, self.n, self.size, self.device = blocks, n, size, device
        self.feature_tap = feature_tap
        self.ops = []          # callables fn(b0, nb) enqueueing one kernel each over frames [b0, b0+nb)
        self.op_kinds = []     # "conv" | "maxpool" | "upsample" | "decode", and the cfg block of every op
        self.op_blocks = []
        # Decode kernels form a second launch list (post_ops): in a stream of batches they run with the NMS on the
        # post-processing stream while the next batch's convolutions run (models.DetectPipeline).
        self.post_ops = []
        self.post_blocks = []
        # ME_FUSE_DECODE=1 runs the YOLO decode in the epilogue of the head conv instead (me_conv_gemm_yolo, bit-identical
        # results).  Measured slower on B200 (batch 32: conv 2.95 ms vs 2.65 ms + 0.09 ms of decode kernels): the eight
        # epilogue warps of a 1-CTA/SM GEMM are instruction-bound on the exp/sigmoid of 255 channels, so it is opt-in.
        self.fuse_decode = os.environ.get("ME_FUSE_DECODE", "0") == "1"
        # Frames are independent, so the batch is run as `splits` sub-batches on parallel streams inside one
        # CUDA graph: while one sub-batch's persistent kernel drains its last (partial) wave of tiles, the
        # other sub-batch's kernel takes over the idle SMs.
        # (measured on B200, profiles/round1/splits_sweep.log: 1 -> 10.4k, 2 -> 10.3-10.5k, 4 -> 9.3k frames/s at
        # batch 32: the persistent kernels already keep the SMs busy, so the default stays at one stream.)
        if splits is None:
            splits = 1
        self.splits = splits if n % splits == 0 else 1
        self._streams = None
        self.graph = None
        self.launches = 0
        self._tensors = tensors
        # Two input buffers (and one captured graph per buffer): the host->device copy of batch i+1 runs on a
        # copy stream while the graph of batch i is still executing.
        self._x_bufs = [torch.zeros((n, in_channels, size, size), dtype=torch.float32, device=device) for _ in range(2)]
        self._u8_bufs = [None, None]     # staging for uint8 frames (load_input)
        self._slot = 0
        self._graphs = [None, None]
        self._post_graphs = [None, None]
        self._conv_ws = []
        self._ws_now = None
        self._conv_meta = {}   # op index -> conv launch description
        self.chains = []       # ops.ConvChain objects (kept alive; the op list holds their run())
        # ME_CONV_CHAIN=0: one kernel per conv layer (per-layer tools: conv_trace.py, ncu_layers.py)
        self.use_chains = os.environ.get("ME_CONV_CHAIN", "1") != "0" and torch.device(device).type == "cuda"
        self._slot_free = [None, None]   # event: last forward that read the buffer has finished
        self._out_busy = [None, None]    # event: a consumer on another stream is done with this slot's yolo_out
        self._copy_stream = None
        self._build()

    # ------------------------------------------------------------------ plan construction
    def _new(self, h, w, c, dtype=torch.float16):
        return torch.zeros((self.n, h, w, c), dtype=dtype, device=self.device)

    def _build(self):
        blocks, n = self.blocks, self.n
        nb = len(blocks)
        # who reads each block's output (besides the next block)
        readers = {i: [] for i in range(nb)}
        for i, b in enumerate(blocks):
            if b["type"] == "route":
                for j in b["layers"]:
                    readers[j].append(i)
            elif b["type"] == "shortcut":
                readers[b["src"]].append(i)
                readers[i - 1].append(i)
        # spatial size of every block output
        hw = []
        s = self.size
        for i, b in enumerate(blocks):
            t = b["type"]
            if t == "convolutional":
                pad = (b["size"] - 1) // 2
                s = (s + 2 * pad - b["size"]) // b["stride"] + 1
            elif t == "maxpool":
                s = s // 2 if b["stride"] == 2 else s
            elif t == "upsample":
                s = s * b["stride"]
            elif t == "route":
                s = hw[b["layers"][0]]
            elif t == "shortcut":
                s = hw[i - 1]
            hw.append(s)
        self.hw = hw

        # two-input routes become channel slices of one concat buffer their producers write into
        target = {}     # block index -> (concat buffer, channel offset)
        for i, b in enumerate(blocks):
            if b["type"] == "route" and len(b["layers"]) > 1:
                total = sum(blocks[j]["out_c"] for j in b["layers"])
                buf = self._new(hw[i], hw[i], total)
                off = 0
                for j in b["layers"]:
                    if j in target:
                        raise MeError("a block feeds two concats; unsupported cfg")
                    if blocks[j]["out_c"] % 32 != 0:
                        raise MeError("concat inputs must have a multiple of 32 channels")
                    target[j] = (buf, off)
                    off += blocks[j]["out_c"]
                b["_concat"] = buf

        # yolo output rows
        self.attrs = None
        rows_total = 0
        for i, b in enumerate(blocks):
            if b["type"] == "yolo":
                rows_total += len(b["anchors"]) * hw[i] * hw[i]
                self.attrs = 5 + b["classes"]
        self.rows_total = rows_total
        # one decoded-output buffer per slot: post-processing of batch i (another stream) overlaps the forward of i+1
        self._yolo_bufs = [torch.zeros((n, rows_total, self.attrs), dtype=torch.float32, device=self.device)
                           for _ in range(2)]

        views = [None] * nb
        self._pooled = {}      # max-pool block index -> view written by the conv in front of it (fused pool)
        row_off = 0
        self.feature_view = None
        for i, b in enumerate(blocks):
            t = b["type"]
            src = views[i - 1] if i > 0 else None
            s_out = hw[i]
            if t == "convolutional":
                fuse_res = None
                nxt = blocks[i + 1] if i + 1 < nb else None
                if nxt is not None and nxt["type"] == "shortcut":
                    if readers[i] != [i + 1]:
                        raise MeError("conv feeding a shortcut is also read elsewhere; unsupported cfg")
                    fuse_res = nxt["src"]
                is_head = nxt is not None and nxt["type"] == "yolo"
                out_idx = i + 1 if fuse_res is not None else i
                cout = b["filters"]
                if is_head and self.fuse_decode and b["size"] == 1 and not b["leaky"] and readers[i] == [] and i not in target:
                    g = s_out
                    packed = self._pack(i, b)
                    self._add(lambda b0, nb, sv=src, p=packed, bb=nxt, g=g, st=self.size / g, ro=row_off: ops.conv_gemm_yolo(
                        sv.at(b0), p, nb, sv.h, sv.w, sv.pitch, self.yolo_out[b0:], g, bb["anchors"], bb["classes"], st,
                        self.rows_total, ro, cin=sv.real_c if sv.real_c != sv.c else sv.c, cout=p.cout_pad), "conv", i)
                    nxt["_decoded_by_conv"] = True
                    views[i] = None
                    continue
                pool_next = (i == 0 and nxt is not None and nxt["type"] == "maxpool" and nxt["size"] == 2 and nxt["stride"] == 2
                             and fuse_res is None and readers[0] == [] and 0 not in target and 1 not in target
                             and b["size"] == 3 and b["stride"] == 1 and b["cin"] <= 3 and cout in (16, 32)
                             and s_out % 32 == 0 and os.environ.get("ME_FUSE_POOL", "1") != "0")
                if pool_next:
                    # blocks 0-1 of the tiny cfgs: the 2x2 max-pool runs in the first conv's epilogue, the full-resolution
                    # tensor (the largest of the network) is never written
                    pv = View(self._new(s_out // 2, s_out // 2, ops.round_up(cout, 8)), 0, ops.round_up(cout, 8),
                              s_out // 2, s_out // 2, real_c=cout)
                    self._add_conv(i, b, src, pv, None, views, False, pooled=True)
                    views[i] = None
                    self._pooled[1] = pv
                    continue
                if (i > 0 and nxt is not None and nxt["type"] == "maxpool" and nxt["size"] == 2 and nxt["stride"] == 2
                        and fuse_res is None and not is_head and readers[i] == [] and i not in target and i + 1 not in target
                        and b["size"] == 3 and b["stride"] == 1 and self.device.type == "cuda"
                        and b["cin"] == 16 and cout in (32, 64) and i != self.feature_tap   # 32-channel inputs: the TMA
                        # kernel + a separate pool is 7 us faster at 104^2 x 32 (profiles/round2/launches_fusion_r2z_summary.txt)
                        and os.environ.get("ME_FUSE_POOL", "1") != "0"):
                    packed = self._pack(i, b)
                    cin_real = src.real_c if src.real_c != src.c else src.c
                    cpad = packed.cout_pad
                    if ops.conv_pool_supported(packed, self.n, src.h, src.w, src.pitch, cpad,
                                               ME_ACT_LEAKY if b["leaky"] else ME_ACT_LINEAR, cin=cin_real, cout=cpad):
                        # thin 3x3 layer + 2x2 max-pool (tiny cfgs, blocks 2-3 / 4-5): pooled in the conv's epilogue
                        pv = View(self._new(s_out // 2, s_out // 2, cpad), 0, cpad, s_out // 2, s_out // 2, real_c=cout)
                        self._add(lambda b0, nb, sv=src, p=packed, o=pv, a=ME_ACT_LEAKY if b["leaky"] else ME_ACT_LINEAR,
                                  ci=cin_real: ops.conv_pool(sv.at(b0), p, nb, sv.h, sv.w, sv.pitch, o.at(b0), o.pitch, act=a,
                                                             cin=ci, cout=p.cout_pad), "conv", i)
                        views[i] = None
                        self._pooled[i + 1] = pv
                        continue
                if out_idx in target:
                    buf, off = target[out_idx]
                    ov = View(buf, off, cout, s_out, s_out)
                else:
                    # the SIMT first conv writes exactly cout channels; the GEMM writes whole 32-wide groups
                    cpad = ops.round_up(cout, 8 if i == 0 else 32)
                    ov = View(self._new(s_out, s_out, cpad, torch.float32 if is_head else torch.float16), 0, cpad,
                              s_out, s_out, real_c=cout)
                    if is_head:   # head logits live per slot, like the decoded output they are turned into
                        ov.alt = View(self._new(s_out, s_out, cpad, torch.float32), 0, cpad, s_out, s_out, real_c=cout)
                self._add_conv(i, b, src, ov, fuse_res, views, is_head)
                views[i] = ov
                if fuse_res is not None:
                    b["_fused_into_next"] = True
            elif t == "shortcut":
                if not blocks[i - 1].get("_fused_into_next"):
                    raise MeError("shortcut without a preceding conv; unsupported cfg")
                views[i] = views[i - 1]
            elif t == "maxpool":
                if i in self._pooled:
                    views[i] = self._pooled[i]         # already produced by the preceding conv's epilogue
                    continue
                if b["size"] != 2:
                    raise MeError("only 2x2 max-pool is supported")
                if i in target:
                    raise MeError("a max-pool that feeds a two-input route is not supported (its output would have to be "
                                  "written into the concat buffer)")
                ov = View(self._new(s_out, s_out, ops.round_up(src.c, 8)), 0, src.c, s_out, s_out, real_c=src.real_c)
                self._add(lambda b0, nb, sv=src, o=ov, st=b["stride"]: ops.maxpool2(
                    sv.at(b0), o.at(b0), nb, sv.h, sv.w, ops.round_up(sv.c, 8), sv.pitch, o.pitch, st), "maxpool", i)
                views[i] = ov
            elif t == "upsample":
                if b["stride"] != 2:
                    raise MeError("only x2 upsample is supported")
                if i in target:
                    buf, off = target[i]
                    ov = View(buf, off, src.c, s_out, s_out, real_c=src.real_c)
                else:
                    ov = View(self._new(s_out, s_out, ops.round_up(src.c, 8)), 0, src.c, s_out, s_out, real_c=src.real_c)
                self._add(lambda b0, nb, sv=src, o=ov: ops.upsample2(sv.at(b0), o.at(b0), nb, sv.h, sv.w, sv.c, sv.pitch,
                                                                     o.pitch), "upsample", i)
                views[i] = ov
            elif t == "route":
                if len(b["layers"]) == 1:
                    views[i] = views[b["layers"][0]]
                else:
                    buf = b["_concat"]
                    views[i] = View(buf, 0, buf.shape[-1], s_out, s_out)
            elif t == "yolo":
                g = s_out
                stride = self.size / g
                if not b.get("_decoded_by_conv"):
                    self._add(lambda b0, nb, sv=src, bb=b, g=g, st=stride, ro=row_off: ops.yolo_decode(
                        self._slot_view(sv).at(b0), sv.pitch, self.yolo_out[b0:], nb, g, bb["anchors"], bb["classes"], st,
                        self.rows_total, ro), "decode", i, post=True)
                row_off += len(b["anchors"]) * g * g
                views[i] = src
            if i == self.feature_tap:
                self.feature_view = views[i]
        self.views = views
        if self.use_chains and self.splits == 1:
            self._form_chains()
        if self.feature_tap is not None and self.feature_view is not None and self.feature_view.buf.dtype != torch.float16:
            raise MeError("feature tap must be an fp16 activation")

    def _form_chains(self):
        """Groups runs of consecutive chain-eligible conv launches into one persistent kernel each (ops.ConvChain).

        Head convs (their output only feeds the decode kernels) are moved to the end of the launch list first, so
        the runs they used to interrupt join up: Darknet-53 becomes 9 per-layer launches, three chains (52^2..13^2
        trunk + first head trunk, 26^2 head trunk, 52^2 head trunk) separated by the two upsamples, and the three
        head convs."""
        n_ops = len(self.ops)

        def describe(k):
            m = self._conv_meta[k]
            sv, o, rv = m["src"], m["out"], m["res"]
            return ops.conv_desc(m["packed"], self.n, sv.h, sv.w, sv.pitch, o.pitch, stride=m["stride"], act=m["act"],
                                 res_pitch=0 if rv is None else rv.pitch, cin=m["cin"], cout=m["cout"], out_f32=m["f32"])

        # Measured on B200 (profiles/round2/chain_trace_r2k.log): the 256-row pair tiles lose to the single-CTA kernels on
        # the thin 104^2 layers (1x1 128 -> 64 etc.: 2 K blocks per tile, the epilogue outlasts the MMAs - the run of
        # blocks 5..9 took 285 us chained against 212 us as four launches), so layers with fewer than 128 filters, and 1x1
        # layers over fewer than 128 channels, stay per-layer launches.
        def profitable(m):
            return m["cout"] >= 128 and (m["cin"] >= 128 or m["packed"].ksize == 3)

        # head convs that cannot join a chain go to the end of the list so that they do not interrupt a run
        def deferred(k):
            m = self._conv_meta.get(k)
            return m is not None and m["leaf"] and not (ops.conv_chain_eligible(describe(k)) and profitable(m))

        order = [k for k in range(n_ops) if not deferred(k)] + [k for k in range(n_ops) if deferred(k)]
        # producer op of every conv output view: buffer -> [(channel offset, channels, op)]
        written = {}
        for k in order:
            m = self._conv_meta.get(k)
            if m is not None:
                o = m["out"]
                written.setdefault(id(o.buf), []).append((o.off, o.c, k))

        def producers(view):
            if view is None:
                return []
            return [k for (off, c, k) in written.get(id(view.buf), []) if off < view.off + view.c and view.off < off + c]

        new_ops, new_kinds, new_blocks = [], [], []
        run = []   # op indices of the chain being collected

        def flush():
            if len(run) >= 2:
                pos = {k: j for j, k in enumerate(run)}
                # head logits exist once per slot (the decode of batch i overlaps the forward of batch i+1): a chain
                # that writes them is built once per slot and the launch picks the current one
                slots = 2 if any(getattr(self._conv_meta[k]["out"], "alt", None) is not None for k in run) else 1
                per_slot = []
                for slot in range(slots):
                    layers = []
                    for k in run:
                        m = self._conv_meta[k]
                        dep = [pos[q] for q in producers(m["src"]) if q in pos]
                        res = [pos[q] for q in producers(m["res"]) if q in pos]
                        o = m["out"]
                        if slot == 1 and getattr(o, "alt", None) is not None:
                            o = o.alt
                        layers.append(dict(desc=describe(k), x=m["src"].t, packed=m["packed"], y=o.t,
                                           residual=None if m["res"] is None else m["res"].t,
                                           dep=dep[0] if dep else -1, res=res[0] if res else -1))
                    per_slot.append(ops.ConvChain(layers, self.device))
                self.chains.extend(per_slot)
                new_ops.append(lambda b0, nb, cs=per_slot: cs[self._slot if len(cs) > 1 else 0].run())
                new_kinds.append("conv")
                new_blocks.append([self.op_blocks[k] for k in run])
            else:
                for k in run:
                    new_ops.append(self.ops[k])
                    new_kinds.append(self.op_kinds[k])
                    new_blocks.append(self.op_blocks[k])
            run.clear()

        # ME_CHAIN_CUT=b1,b2,...: start a new chain in front of these cfg blocks (stage-level timing, tools/chain_trace.py)
        cuts = {int(v) for v in os.environ.get("ME_CHAIN_CUT", "").split(",") if v}
        for k in order:
            m = self._conv_meta.get(k)
            ok = m is not None and ops.conv_chain_eligible(describe(k)) and profitable(m)
            if ok and self.op_blocks[k] in cuts:
                flush()
            if ok:
                in_run = set(run)
                # at most one producer inside the run for the input and for the residual (a concat of two in-run
                # layers would need two dependencies)
                if len([q for q in producers(m["src"]) if q in in_run]) > 1 or \
                        len([q for q in producers(m["res"]) if q in in_run]) > 1:
                    flush()
                run.append(k)
            else:
                flush()
                new_ops.append(self.ops[k])
                new_kinds.append(self.op_kinds[k])
                new_blocks.append(self.op_blocks[k])
        flush()
        self.ops, self.op_kinds, self.op_blocks = new_ops, new_kinds, new_blocks

    def _bn_of(self, i):
        p = f"module_list.{i}.batch_norm_{i}."
        t = self._tensors
        return (t[p + "weight"], t[p + "bias"], t[p + "running_mean"], t[p + "running_var"], 1e-5)

    def _add(self, fn, kind, block, post=False):
        if post:
            self.post_ops.append(fn)
            self.post_blocks.append(block)
            return
        self.ops.append(fn)
        self.op_kinds.append(kind)
        self.op_blocks.append(block)

    def head_logits(self):
        """[(fp32 logits view of the current slot, yolo block, grid size)] in cfg order (empty views are skipped when
        the decode is fused into the head conv)."""
        out = []
        for i, b in enumerate(self.blocks):
            if b["type"] == "yolo" and self.views[i] is not None:
                out.append((self._slot_view(self.views[i]), b, self.hw[i]))
        return out

    def _slot_view(self, v):
        """The view's buffer of the current slot (only head-logit views have a second one)."""
        alt = getattr(v, "alt", None)
        return alt if (alt is not None and self._slot == 1) else v

    def _pack(self, i, b):
        t = self._tensors
        packed = ops.pack_conv(t[f"module_list.{i}.conv_{i}.weight"], t.get(f"module_list.{i}.conv_{i}.bias"),
                               self._bn_of(i) if b["bn"] else None, cout_pad=ops.round_up(b["filters"], 32))
        self._keep = getattr(self, "_keep", []) + [packed]
        return packed

    def _add_conv(self, i, b, src, ov, fuse_res, views, is_head, pooled=False):
        n = self.n
        t = self._tensors
        w = t[f"module_list.{i}.conv_{i}.weight"]
        bias = t.get(f"module_list.{i}.conv_{i}.bias")
        bn = self._bn_of(i) if b["bn"] else None
        act = ME_ACT_LEAKY if b["leaky"] else ME_ACT_LINEAR
        if i == 0:
            if b["size"] != 3 or b["stride"] != 1 or b["cin"] > 4:
                raise MeError("first layer must be a 3x3/stride-1 conv over <= 4 input channels")
            first = ops.pack_first_conv(w, bias, bn)
            self._keep = getattr(self, "_keep", []) + [first]
            self._add(lambda b0, nb, f=first, o=ov, a=act, pl=pooled: ops.conv_first(self.x_in[b0:b0 + nb], f, o.at(b0), o.pitch, a,
                                                                                      pool=pl), "conv", i)
            return
        packed = self._pack(i, b)
        res_v = views[fuse_res] if fuse_res is not None else None
        # what the chain builder (_form_chains) needs to describe this launch to me_conv_chain_build
        self._conv_meta[len(self.ops)] = dict(src=src, packed=packed, out=ov, stride=b["stride"], act=act, res=res_v,
                                              cin=src.real_c if src.real_c != src.c else src.c, cout=packed.cout_pad,
                                              f32=is_head, leaf=is_head)
        self._add(lambda b0, nb, sv=src, p=packed, o=ov, st=b["stride"], a=act, rv=res_v, f32=is_head: ops.conv_gemm(
            sv.at(b0), p, nb, sv.h, sv.w, sv.pitch, self._slot_view(o).at(b0), o.pitch, stride=st, act=a,
            residual=None if rv is None else rv.at(b0), res_pitch=0 if rv is None else rv.pitch,
            cin=sv.real_c if sv.real_c != sv.c else sv.c, cout=p.cout_pad, out_f32=f32, workspace=self._ws_now), "conv", i)

    # ------------------------------------------------------------------ execution
    def enqueue(self, b0=0, nb=None, only=None):
        nb = self.n if nb is None else nb
        # one split-K workspace per sub-batch stream: the conv kernels of different sub-batches may overlap
        k = b0 // nb if nb else 0
        while len(self._conv_ws) <= k:
            self._conv_ws.append(ops.conv_workspace(self.device))
        self._ws_now = self._conv_ws[k]      # what the conv ops of this sub-batch pass to me_conv_gemm_ws
        for i, fn in enumerate(self.ops):
            if only is None or i in only:
                fn(b0, nb)

    def enqueue_post(self):
        """The decode kernels of the current slot (head logits -> yolo_out)."""
        for fn in self.post_ops:
            fn(0, self.n)

    def enqueue_split(self, only=None):
        """All sub-batches, each on its own stream, joined back into the current stream.
        `only`: optional set of op indices (profiling a subset of the launch list)."""
        if self.splits == 1:
            self.enqueue(only=only)
            self.launches = len(self.ops) + len(self.post_ops)
            return
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.splits - 1)]
        cur = torch.cuda.current_stream()
        nb = self.n // self.splits
        for s in self._streams:
            s.wait_stream(cur)
        self.enqueue(0, nb, only)
        for k, s in enumerate(self._streams):
            with torch.cuda.stream(s):
                self.enqueue((k + 1) * nb, nb, only)
        for s in self._streams:
            cur.wait_stream(s)
        self.launches = len(self.ops) * self.splits + len(self.post_ops)

    @property
    def yolo_out(self):
        """Decoded predictions of the current slot (written by the last run())."""
        return self._yolo_bufs[self._slot]

    def hold_output(self, event):
        """A consumer running on another stream records `event` when it no longer needs the current slot's
        yolo_out; the next run() that writes this slot waits for it."""
        self._out_busy[self._slot] = event

    @property
    def x_in(self):
        """Input buffer of the current slot (what the first conv reads)."""
        return self._x_bufs[self._slot]

    def next_input(self):
        """The (n, C, S, S) fp32 device buffer the NEXT forward reads.  A producer that already works on the device
        (decoder, augmentation, a previous model) writes the batch straight into it and passes this very tensor to
        forward_device(): no staging copy is made."""
        return self._x_bufs[self._slot ^ 1]

    def load_input(self, x):
        """Stages the next batch: switches to the other input buffer and copies x into it (nothing to copy when x is
        next_input()).  A pinned host tensor is copied on a separate stream, so the PCIe transfer overlaps the previous
        forward.  x may be fp32 in 0..1 (the reference's contract) or uint8 0..255 in the same (N,C,S,S) layout: bytes
        are uploaded as they are and divided by 255 on the device - exactly ToTensor's values, a quarter of the traffic."""
        self._slot ^= 1
        buf = self._x_bufs[self._slot]
        cur = torch.cuda.current_stream()
        is_u8 = x.dtype == torch.uint8
        if is_u8 and (tuple(x.shape) != tuple(buf.shape) or not x.is_contiguous()):
            raise MeError(f"uint8 frames must be contiguous {tuple(buf.shape)} (N,C,S,S) bytes")
        if is_u8 and self._u8_bufs[self._slot] is None:
            self._u8_bufs[self._slot] = torch.zeros(buf.shape, dtype=torch.uint8, device=self.device)
        if x.is_cuda:
            if is_u8:
                ops.u8_to_unit_f32(x, buf)
            elif x.data_ptr() != buf.data_ptr() or x.shape != buf.shape or not x.is_contiguous():
                buf.copy_(x, non_blocking=True)
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        if self._slot_free[self._slot] is not None:
            cs.wait_event(self._slot_free[self._slot])
        with torch.cuda.stream(cs):
            if is_u8:
                # camera frames as bytes: a quarter of the PCIe traffic; ToTensor's x / 255 runs on the device
                self._u8_bufs[self._slot].copy_(x, non_blocking=True)
                ops.u8_to_unit_f32(self._u8_bufs[self._slot], buf)
            else:
                buf.copy_(x, non_blocking=True)
        cur.wait_stream(cs)

    def run_decode(self, use_graph=True):
        """Decode kernels of the current slot on the current stream (after run(decode=False))."""
        if not self.post_ops:
            return
        if not use_graph:
            self.enqueue_post()
            return
        if self._post_graphs[self._slot] is None:
            self.enqueue_post()
            torch.cuda.synchronize(self.device)
            self._post_graphs[self._slot] = capture_graph(self.enqueue_post)
        self._post_graphs[self._slot].replay()

    def run(self, use_graph=True, decode=True):
        """Inputs must already be staged with load_input(). Enqueues (or replays) the whole forward; decode=False
        leaves the head logits undecoded for a later run_decode() (possibly on another stream)."""
        busy = self._out_busy[self._slot]
        if busy is not None:
            torch.cuda.current_stream().wait_event(busy)
            self._out_busy[self._slot] = None
        if not use_graph:
            self.enqueue_split()
        else:
            if self._graphs[self._slot] is None:
                self.enqueue_split()  # warm-up: lazy one-time initialisation inside the library
                torch.cuda.synchronize(self.device)
                self._graphs[self._slot] = capture_graph(self.enqueue_split)
            self.graph = self._graphs[self._slot]
            self.graph.replay()
        ev = self._slot_free[self._slot]
        if ev is None:
            ev = self._slot_free[self._slot] = torch.cuda.Event()
        ev.record()
        if decode:
            self.run_decode(use_graph)
