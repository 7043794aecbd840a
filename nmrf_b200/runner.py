"""CUDA-graph runner: the whole forward (feature extractor + hot path, every launch a libnmrf_b200 kernel) captured once
per input shape and replayed, so a step costs one graph launch instead of ~200 kernel launches."""
import torch


class GraphedNMRF:
    """`runner(img1, img2)`: images may live on the host (pinned memory recommended) or on the device."""

    def __init__(self, model, B, H, W, warmup=3, graph_full=True):
        """graph_full=False: only the hot path is captured; the feature extractor runs eagerly.  For foreign encoders whose
        forward cannot be captured (the reference's DeformNeck builds CPU tensors per call, adaptor_modules.py:27-34)."""
        assert model.device.type == "cuda"
        self.model = model
        dev = model.device
        self.img1 = torch.zeros(B, 3, H, W, device=dev)
        self.img2 = torch.zeros(B, 3, H, W, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # builds the plan, sets kernel attributes, warms cuDNN
                model.forward_device(self.img1, self.img2)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = None
        if graph_full:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = model.forward_device(self.img1, self.img2)
        else:
            self.out = model.forward_device(self.img1, self.img2)
        self.plan = model.plan_for(B, *self._feat_shape(model, B, H, W), H, W)
        self.disp_host = torch.empty(B, H, W, pin_memory=True)
        # the hot path alone (libnmrf_b200 kernels only; its inputs -- the feature maps -- stay resident in the plan)
        self.hot_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.hot_graph):
            self.plan.run()

    @staticmethod
    def _feat_shape(model, B, H, W):
        d = model.divis_by
        Hp, Wp = H + (((H // d) + 1) * d - H) % d, W + (((W // d) + 1) * d - W) % d
        enc = model.backbone if model.compat else model.image_encoder
        return enc.output_dim, Hp // 8, Wp // 8

    def replay(self):
        """inputs already in self.img1/img2 (device-resident step)"""
        if self.graph is not None:
            self.graph.replay()
        else:
            self.out = self.model.forward_device(self.img1, self.img2)
        return self.out

    def replay_hot_path(self):
        """one pass of the hot path over the feature maps currently resident in the plan"""
        self.hot_graph.replay()
        return self.plan.disp

    # ---- streaming API: host pairs in, host disparities out, copies overlapped with the previous / next pair's compute ----
    def _pipeline_init(self):
        dev = self.img1.device
        self._copy = torch.cuda.Stream(device=dev)
        self._stage_in = [(torch.empty_like(self.img1), torch.empty_like(self.img2)) for _ in range(2)]
        self._stage_out = [torch.empty_like(self.out["disp"]) for _ in range(2)]
        # THREE pinned result buffers: result i lands in buffer i % 3, so the buffer handed out for result i - 1 is not
        # written again before result i + 2 is enqueued, i.e. before the consumer has asked for the next-but-one item
        self._host_out = [torch.empty(self.out["disp"].shape, pin_memory=True) for _ in range(3)]
        ev = lambda: torch.cuda.Event()
        self._ev_h2d, self._ev_in_free, self._ev_done, self._ev_d2h = ([ev(), ev()] for _ in range(4))
        self._ev_host = [ev() for _ in range(3)]
        self._seq = 0
        self._primed = False

    def _prefetch(self, slot, img1, img2):
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(self._ev_in_free[slot])          # the d2d copy that last read this staging pair is done
            self._stage_in[slot][0].copy_(img1, non_blocking=True)
            self._stage_in[slot][1].copy_(img2, non_blocking=True)
            self._ev_h2d[slot].record(self._copy)

    def stream(self, pairs):
        """Generator over (img1, img2) HOST pairs (pinned memory): yields each pair's disparity as a pinned host tensor that stays
        valid until the next-but-one result has been requested (three result buffers rotate).  The H2D copy of pair i+1 and the D2H copy of pair i-1 run on a copy
        stream while pair i computes (the graph itself is unchanged: staging buffers are copied device-to-device)."""
        if not hasattr(self, "_copy"):
            self._pipeline_init()
        main = torch.cuda.current_stream(self.img1.device)
        it = iter(pairs)
        nxt = next(it, None)
        if nxt is None:
            return
        for e in self._ev_in_free + self._ev_d2h:
            e.record(main)
        self._prefetch(self._seq & 1, *nxt)
        pending = None
        while nxt is not None:
            slot = self._seq & 1
            cur, nxt = nxt, next(it, None)
            if nxt is not None:
                self._prefetch(slot ^ 1, *nxt)
            main.wait_event(self._ev_h2d[slot])
            self.img1.copy_(self._stage_in[slot][0]); self.img2.copy_(self._stage_in[slot][1])
            self._ev_in_free[slot].record(main)
            self.replay()
            main.wait_event(self._ev_d2h[slot])                      # the D2H that last read this output staging buffer is done
            self._stage_out[slot].copy_(self.out["disp"])
            self._ev_done[slot].record(main)
            hs = self._seq % 3
            with torch.cuda.stream(self._copy):
                self._copy.wait_event(self._ev_done[slot])
                self._host_out[hs].copy_(self._stage_out[slot], non_blocking=True)
                self._ev_d2h[slot].record(self._copy)
                self._ev_host[hs].record(self._copy)
            if pending is not None:
                self._ev_host[pending].synchronize()
                yield self._host_out[pending]
            pending = hs
            self._seq += 1
        self._ev_host[pending].synchronize()
        yield self._host_out[pending]

    def __call__(self, img1, img2, to_host=False):
        self.img1.copy_(img1, non_blocking=True)
        self.img2.copy_(img2, non_blocking=True)
        self.replay()
        if to_host:
            self.disp_host.copy_(self.out["disp"], non_blocking=True)
            return self.disp_host
        return self.out
