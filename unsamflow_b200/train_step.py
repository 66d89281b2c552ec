"""Synthetic-input PWC-Lite training step, single GPU or DDP (one process per GPU).

Mirrors what the reference does per iteration at stage 0 (trainer/kitti_trainer_ar.py:99-131,309-317):
model(img1, img2, with_bk=True) -> cat(flow12, flow21) per scale -> unFlowLoss -> zero_grad ->
backward -> clip_grad_norm_(10) -> Adam step, and how it launches (train.py:42-147): NCCL process
group, per-GPU batch = batch_size // world_size, DistributedDataParallel on the flow model only (the
reference also wraps an unused MaskUNet, train.py:121-126 -- not replicated).
The 10 MB gradient all-reduce is the only collective; DDP is configured for one bucket
(bucket_cap_mb=16, gradient_as_bucket_view, static_graph) as SURVEY.md 5 recommends.
"""
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel

from .flow_loss import UnFlowLoss
from .pwclite import PWCLite


class TrainStep:
    def __init__(self, batch, height, width, device, seed=42, batch_directions=False, ddp=False, lr=2e-4,
                 weight_decay=1e-6, bias_decay=0.0, cuda_graph=False, fuse_cost_concat=False):
        """cuda_graph=True (single GPU): the whole iteration -- forward, loss, backward, clipping, Adam -- is captured
        once into a CUDA graph and replayed (static input buffers, capturable Adam); the usf ops never synchronise
        or touch the host, so nothing in the step prevents it."""
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.model = PWCLite(batch_directions=batch_directions, fuse_cost_concat=fuse_cost_concat).to(self.device)
        self.model.init_weights()
        self.net = self.model
        if ddp:
            self.net = DistributedDataParallel(
                self.model, device_ids=[self.device.index] if self.device.type == "cuda" else None,
                bucket_cap_mb=16, gradient_as_bucket_view=True, static_graph=True)
        self.loss_fn = UnFlowLoss()
        # trainer/base_trainer.py:140-165 with configs/kitti_base.json: Adam(lr 2e-4, betas (0.9, 0.999), eps 1e-7)
        # over two parameter groups -- weights with weight_decay 1e-6, biases with bias_decay 0
        # (utils/torch_utils.py:27-41: ".weight" / ".bias" in the parameter name, everything else undecayed)
        named = list(self.model.named_parameters())
        weights = [p for n, p in named if ".weight" in n]
        biases = [p for n, p in named if ".bias" in n]
        others = [p for n, p in named if ".weight" not in n and ".bias" not in n]
        groups = [{"params": biases, "weight_decay": bias_decay}, {"params": weights, "weight_decay": weight_decay}]
        if others:
            groups.append({"params": others, "weight_decay": 0.0})
        if cuda_graph and (ddp or self.device.type != "cuda"):
            raise ValueError("cuda_graph=True is the single-GPU form of the step (DDP's bucket hooks are not captured)")
        self.opt = torch.optim.Adam(groups, lr=lr, betas=(0.9, 0.999), eps=1e-7, capturable=bool(cuda_graph))
        self.cuda_graph, self.graph, self.static_loss = cuda_graph, None, None
        gen = torch.Generator().manual_seed(seed + (dist.get_rank() if ddp else 0))
        self.img1 = torch.rand(batch, 3, height, width, generator=gen).to(self.device)
        self.img2 = torch.rand(batch, 3, height, width, generator=gen).to(self.device)
        self.batch = batch

    def step(self, img1=None, img2=None):
        if self.cuda_graph:
            return self._graph_step(img1, img2)
        img1 = self.img1 if img1 is None else img1
        img2 = self.img2 if img2 is None else img2
        return self._eager_step(img1, img2)

    def _graph_step(self, img1, img2):
        if img1 is not None:
            self.img1.copy_(img1)
        if img2 is not None:
            self.img2.copy_(img2)
        if self.graph is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):  # cuDNN autotuning, lazy initialisations and the allocator's warm-up stay outside
                    self._eager_step(self.img1, self.img2)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self.static_loss = self._eager_step(self.img1, self.img2)
        self.graph.replay()
        return self.static_loss

    def _eager_step(self, img1, img2):
        out = self.net(img1, img2, with_bk=True)
        flows = [torch.cat([f12, f21], 1) for f12, f21 in zip(out["flows_12"], out["flows_21"])]
        loss = self.loss_fn(flows, img1, img2)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 10.0)
        self.opt.step()
        return loss.detach()
