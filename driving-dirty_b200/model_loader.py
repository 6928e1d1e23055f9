"""ModelLoader: the competition-style inference front-end named by the north star.

The reference tree does not contain it; the interface follows the public NYU DLSP20
``model_loader.py`` template.  ``get_binary_road_map`` is DEFINED by composition of reference code:
``RoadMapBCE.forward(samples)[1].round()`` (roadmap_bce_v2.py:66-81,140), and that is what the
parity tests pin.  ``get_bounding_boxes`` has no importable reference implementation:
the interface is kept, its result is parity-unpinned.
"""
from argparse import Namespace

import torch

from . import ops
from .lightning_compat import save_checkpoint  # noqa: F401
from .roadmap_model.roadmap_bce_v2 import RoadMapBCE


def get_transform_task1():
    import torchvision
    return torchvision.transforms.ToTensor()


def get_transform_task2():
    import torchvision
    return torchvision.transforms.ToTensor()


def get_transform_bytes():
    """Front-end for raw camera bytes: the image stays uint8 [3,H,W]; the /255 of ToTensor
    (data_helper.py:109-114) is folded into the first conv's loads on the device -- 4x fewer host-to-device bytes."""
    import torchvision
    return torchvision.transforms.PILToTensor()


class ModelLoader:
    team_name = "driving-dirty-b200"
    team_number = 0
    round_number = 1
    team_member = []
    contact_email = ""

    def __init__(self, model_file="roadmap_bce.ckpt", device="cuda:0", graph_max_batch=8, pipeline_chunk=32):
        """``model_file``: a ``{'state_dict', 'hparams'}`` checkpoint of RoadMapBCE (hparams must
        carry ``pretrained_path`` of the AE checkpoint, as the reference's constructor needs it),
        or an already constructed RoadMapBCE."""
        if isinstance(model_file, RoadMapBCE):
            model = model_file
        else:
            ckpt = torch.load(model_file, map_location="cpu", weights_only=False)
            hp = ckpt["hparams"]
            model = RoadMapBCE(hp if isinstance(hp, Namespace) else Namespace(**hp))
            model.load_state_dict(ckpt["state_dict"])
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.model.eval()
        self._staging = {}          # (shape, dtype) -> [pinned host buffer, device buffer, copy-done event]
        self._copy_stream = None
        # small batches are launch-bound (~30 launches, 0.5 ms at B = 1): up to this batch size the forward is captured
        # once per input shape into a CUDA graph and replayed (0 disables)
        self.graph_max_batch = int(graph_max_batch)
        # page-locked host batches larger than this are copied in chunks of this many scenes, overlapped with the conv stack
        # (0 disables): at batch 256 the 361 MB of camera bytes take as long to cross PCIe as the whole forward pass
        self.pipeline_chunk = int(pipeline_chunk)
        self._graphs = {}           # (shape, dtype) -> (graph, static input, static output)

    def stage(self, samples):
        """Host samples ([B,6,3,H,W] uint8 bytes or fp32) -> device, through a cached PINNED buffer and an asynchronous
        copy on a side stream (a pageable source would make the copy synchronous and staged by the driver)."""
        if samples.is_cuda:
            return samples.to(self.device)
        if samples.is_pinned():                                  # already page-locked: one asynchronous copy, no staging
            return samples.to(self.device, non_blocking=True)
        key = (tuple(samples.shape), samples.dtype)
        slot = self._staging.get(key)
        if slot is None:
            slot = [torch.empty(samples.shape, dtype=samples.dtype).pin_memory(),
                    torch.empty(samples.shape, dtype=samples.dtype, device=self.device), torch.cuda.Event()]
            self._staging[key] = slot
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        host, dev, done = slot
        done.synchronize()                                   # the previous copy out of the pinned buffer has finished
        host.copy_(samples)
        cur = torch.cuda.current_stream(self.device)
        self._copy_stream.wait_stream(cur)                   # the device buffer's previous consumer has finished
        with torch.cuda.stream(self._copy_stream):
            dev.copy_(host, non_blocking=True)
            done.record(self._copy_stream)
        cur.wait_event(done)
        return dev

    @torch.no_grad()
    def get_binary_road_map(self, samples, as_bytes=False):
        """samples: tensor [B,6,3,256,306], fp32 in [0,1] (the competition's CUDA tensor) or uint8 raw camera bytes, on the
        device or on the host (then staged through pinned memory) -> CUDA float tensor [B,800,800] of 0./1., equal to
        ``sigmoid(logits).round()`` of the reference forward (on ``bytes.float() / 255`` for raw bytes).
        ``as_bytes``: the same map as uint8 (what the kernel writes: a quarter of the bytes to store or copy back)."""
        if not samples.is_cuda and samples.is_pinned() and samples.shape[0] > self.pipeline_chunk > 0:
            binary = self._forward_pipelined(samples)
        else:
            x = self.stage(samples)
            binary = self._replay(x) if 0 < x.shape[0] <= self.graph_max_batch else self._forward(x)
        return binary if as_bytes else binary.float()

    def _forward_pipelined(self, samples):
        """Large page-locked host batch: chunk i+1 crosses PCIe (copy stream) while the conv stack runs on chunk i.  The
        conv stack is per scene, so the pooled features are those of one big call; the dense tail -- with its dropout
        draw, which depends on the batch shape -- then runs ONCE over the whole batch: same bits as the one-copy path."""
        enc, c, B = self.model.ae.encoder, self.pipeline_chunk, samples.shape[0]
        cur = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        cs.wait_stream(cur)
        staged = []
        with torch.cuda.stream(cs):
            for i in range(0, B, c):
                x = samples[i:i + c].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                staged.append((x, ev))
        feats = None
        for k, (x, ev) in enumerate(staged):
            cur.wait_event(ev)
            x.record_stream(cur)                     # allocated on the copy stream, consumed here
            f = enc._stack(ops.as_view_batch(x, keep_bytes=True))
            if feats is None:
                feats = torch.empty(B, f.shape[1], dtype=f.dtype, device=self.device)
            feats[k * c:k * c + f.shape[0]] = f
        return ops.binary_map(self.model._head(enc._tail(feats)))

    def _forward(self, x):
        logits = self.model._logits(x)
        return ops.binary_map(logits)

    def _replay(self, x):
        """CUDA-graph path: the whole forward (ToTensor pass, convs, pool, dense blocks with their torch dropout -- the
        generator's Philox offset is graph-safe --, head, sigmoid / binarise) is one graph launch.  Kernel arguments, TMA
        descriptors included, are baked at capture: the input lives in a static buffer the samples are copied into."""
        key = (tuple(x.shape), x.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            rng = torch.cuda.get_rng_state(self.device)     # warm-up and capture must not advance the dropout stream
            static_in = torch.empty_like(x)
            static_in.copy_(x)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                  # warm-up outside capture: lazy module / attribute setup, allocator pools
                for _ in range(2):
                    self._forward(static_in)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._forward(static_in)
            entry = (graph, static_in, static_out)
            self._graphs[key] = entry
            torch.cuda.set_rng_state(rng, self.device)
        graph, static_in, static_out = entry
        static_in.copy_(x)
        graph.replay()
        return static_out.clone()

    @torch.no_grad()
    def get_bounding_boxes(self, samples):
        """Tuple of B tensors [N,2,4] (metres; rows x/y; cols fl, fr, bl, br).  No reference model
        for boxes is importable; one placeholder box per sample keeps the scoring harness alive."""
        B = samples.shape[0]
        box = torch.tensor([[[2.3, 2.3, -2.3, -2.3], [1.0, -1.0, 1.0, -1.0]]], device=self.device)
        return tuple(box.clone() for _ in range(B))
