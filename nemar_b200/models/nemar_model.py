"""NEMARModel on the B200 engine — the training hot path (reference models/nemar_model.py:12-288).

Same plugin surface: flags, loss/visual/model names, set_input / forward / optimize_parameters, netT / netR /
netD attributes and checkpoint keys.  Differences that are part of the design:
  * one process per GPU; gradients of each optimizer phase are exchanged with ONE all-reduce on a flat bucket
    (D bucket after backward_D, T+R bucket after backward_T_and_R — the reference's update order needs two);
  * the three torch.optim.Adam instances become two flat-buffer Adam launches (D, and T+R);
  * real_A resizes for the multi-resolution discriminators are computed once per step, not five times.
"""
import itertools
import os

import torch

from . import networks
from . import stn
from .base_model import BaseModel
from ..engine import functional as F
from ..engine import lib as L
from ..engine import parallel
from ..engine.config import configure
from ..engine.optim import FlatAdam


class NEMARModel(BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        parser = stn.modify_commandline_options(parser, is_train)
        parser.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"],
                            help="[engine] activation/weight storage dtype (accumulation is always fp32)")
        parser.add_argument("--conv_engine", type=str, default="auto", choices=["auto", "generic"],
                            help="[engine] auto: tcgen05 where supported; generic: CUDA-core kernels only")
        parser.add_argument("--cuda_graph", type=int, default=0,
                            help="[engine] 1: capture optimize_parameters (forward, both backward passes, the gradient "
                                 "all-reduces and both Adam launches) in ONE CUDA graph after 3 eager steps and replay it; "
                                 "inputs are copied into static buffers by set_input")
        parser.add_argument("--stream_overlap", type=int, default=1,
                            help="[engine] 1 (default): the registration network's regressor runs on a second CUDA stream, "
                                 "concurrently with netT(real_A) (forward and backward)")
        parser.add_argument("--batch_d", type=int, default=1,
                            help="[engine] 1 (default): the discriminator evaluates its (A, B_k) pairs of one phase (real / fake_TR / "
                                 "fake_RT) in ONE pass over their batch-concatenation instead of one pass each "
                                 "(InstanceNorm keeps samples independent: same losses and gradients, a third of the launches)")
        if is_train:
            parser.add_argument("--lambda_GAN", type=float, default=1.0, help="weight of the GAN loss")
            parser.add_argument("--lambda_recon", type=float, default=100.0, help="weight of the L1 reconstruction loss")
            parser.add_argument("--lambda_smooth", type=float, default=0.0, help="weight of the STN regulariser")
            parser.add_argument("--enable_tbvis", action="store_true", help="tensorboard visualizer (not part of the engine)")
            parser.add_argument("--multi_resolution", type=int, default=1, help="number of discriminator scales")
        return parser

    def __init__(self, opt):
        BaseModel.__init__(self, opt)
        configure(getattr(opt, "precision", "bf16"), getattr(opt, "conv_engine", "auto"))
        self.train_stn = True
        self.setup_visualizers()
        self.tb_visualizer = None        # train.py reads this attribute (reference train.py:77-81)
        self.define_networks()
        if self.isTrain:
            self.criterionGAN = networks.GANLoss(opt.gan_mode).to(self.device)
            self.setup_optimizers()

    def setup_visualizers(self):
        self.loss_names = ["L1_TR", "GAN_TR", "L1_RT", "GAN_RT", "smoothness", "D_fake_TR", "D_fake_RT", "D"]
        self.visual_names = ["real_A", "real_B", "fake_TR_B", "fake_RT_B", "registered_real_A", "fake_B"]
        self.model_names = ["T", "R"] + (["D"] if self.isTrain else [])

    def define_networks(self):
        opt = self.opt
        AtoB = opt.direction == "AtoB"
        in_c = opt.input_nc if AtoB else opt.output_nc
        out_c = opt.output_nc if AtoB else opt.input_nc
        self.netT = networks.define_G(in_c, out_c, opt.ngf, opt.netG, opt.norm, not opt.no_dropout, opt.init_type,
                                      opt.init_gain, self.gpu_ids)
        self.netR = stn.define_stn(opt, opt.stn_type)
        if self.isTrain:
            mk = lambda: networks.define_D(opt.output_nc + opt.input_nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm,
                                           opt.init_type, opt.init_gain, self.gpu_ids)
            self.netD = mk()
            self.netD_multiresolution = [mk() for _ in range(max(opt.multi_resolution - 1, 0))]

    def reset_weights(self):
        opt = self.opt
        for net in [self.netT, self.netD, *self.netD_multiresolution]:
            networks.init_weights(net, opt.init_type, opt.init_gain)

    def setup_optimizers(self):
        opt = self.opt
        betas = (opt.beta1, 0.999)
        # R and T are always stepped together (reference :282-283) -> one flat bucket, one launch
        self.optimizer_TR = FlatAdam(itertools.chain(self.netR.parameters(), self.netT.parameters()), lr=opt.lr, betas=betas)
        d_params = itertools.chain(self.netD.parameters(), *[x.parameters() for x in self.netD_multiresolution])
        self.optimizer_D = FlatAdam(d_params, lr=opt.lr, betas=betas)
        self.optimizer_R = self.optimizer_T = self.optimizer_TR      # reference attribute names
        self.optimizers += [self.optimizer_TR, self.optimizer_D]
        hook = parallel.BucketAllReduce()
        self.optimizer_TR.grad_hook = hook
        self.optimizer_D.grad_hook = hook
        self.allreduce = hook
        # replicas must start identical whatever each process's RNG did before this point
        parallel.broadcast_buffers([self.optimizer_TR.flat_p, self.optimizer_D.flat_p])

    # -- checkpoints: the reference's three files, plus what it forgets -------------------------------------------
    def save_networks(self, epoch):
        """<epoch>_net_{T,R,D}.pth exactly as the reference writes them (base_model.py:148-164), plus two files a
        reference loader never looks for: the extra discriminators of --multi_resolution (plain list in the reference,
        nemar_model.py:108-113, hence never saved there) and the state of both Adam optimizers."""
        BaseModel.save_networks(self, epoch)
        if not self.isTrain:
            return
        import collections
        for i, net in enumerate(self.netD_multiresolution):
            sd = collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items())
            torch.save(sd, os.path.join(self.save_dir, "%s_net_D_ms%d.pth" % (epoch, i + 1)))
        torch.save({"TR": self.optimizer_TR.state_dict(), "D": self.optimizer_D.state_dict()},
                   os.path.join(self.save_dir, "%s_optim.pth" % epoch))

    def load_networks(self, epoch):
        BaseModel.load_networks(self, epoch)
        if not self.isTrain:
            return
        for i, net in enumerate(self.netD_multiresolution):
            path = os.path.join(self.save_dir, "%s_net_D_ms%d.pth" % (epoch, i + 1))
            if os.path.exists(path):          # absent in checkpoints written by the reference
                net.load_state_dict(torch.load(path, map_location="cpu"))
        path = os.path.join(self.save_dir, "%s_optim.pth" % epoch)
        if os.path.exists(path):
            sd = torch.load(path, map_location="cpu")
            self.optimizer_TR.load_state_dict(sd["TR"])
            self.optimizer_D.load_state_dict(sd["D"])
        F.bump_weights_epoch()
        parallel.broadcast_buffers([self.optimizer_TR.flat_p, self.optimizer_D.flat_p])

    def set_input(self, input):
        AtoB = self.opt.direction == "AtoB"
        a, b = ("A", "B") if AtoB else ("B", "A")
        ready = input.get("_ready") if isinstance(input, dict) else None
        if ready is not None:            # staged by data.prefetch.DevicePrefetcher on a side stream
            torch.cuda.current_stream().wait_event(ready)
        if getattr(self, "_graph_inputs", None) is not None and input[a].shape == self._graph_inputs[0].shape:
            # graph mode: the captured kernels read these two static buffers
            self._graph_inputs[0].copy_(input[a], non_blocking=True)
            self._graph_inputs[1].copy_(input[b], non_blocking=True)
            self.real_A, self.real_B = self._graph_inputs
        else:
            self.real_A = input[a].to(self.device, non_blocking=True).float().contiguous()
            self.real_B = input[b].to(self.device, non_blocking=True).float().contiguous()
        self.image_paths = input.get(a + "_paths", [])

    def _stn_stream(self):
        """Second stream for the registration network (None: disabled).  netT(real_A) and the STN's regressor are
        independent in both directions, so the STN's many small, latency-bound kernels can fill the SMs the generator's
        convolutions leave idle (wave tails, one-CTA-per-SM tiles); autograd replays each node on its forward stream, so
        the backward passes overlap the same way.  Works inside the captured step (fork / join become graph edges)."""
        if self.device.type != "cuda" or not getattr(self.opt, "stream_overlap", 1):
            return None
        s = self.__dict__.get("_stn_stream_obj")
        if s is None:
            s = self._stn_stream_obj = torch.cuda.Stream(self.device)
        return s

    def _join_stn_stream(self):
        s = self.__dict__.get("_stn_stream_obj")
        if s is not None:
            torch.cuda.current_stream().wait_stream(s)

    def forward(self):
        side = self._stn_stream() if hasattr(self.netR, "precompute") else None
        if side is not None:
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                pre = self.netR.precompute(self.real_A, self.real_B)
            self.fake_B = self.netT(self.real_A)
            main.wait_stream(side)
            warped, reg_term = self.netR(self.real_A, self.real_B, apply_on=[self.real_A, self.fake_B], pre=pre)
        else:
            self.fake_B = self.netT(self.real_A)
            warped, reg_term = self.netR(self.real_A, self.real_B, apply_on=[self.real_A, self.fake_B])
        self.stn_reg_term = reg_term
        self.registered_real_A = warped[0]
        self.fake_TR_B = self.netT(self.registered_real_A)   # registration first, then translation
        self.fake_RT_B = warped[1]                           # translation first, then registration

    # -- discriminator plumbing --------------------------------------------------------------------
    def _scales(self):
        return [self.netD] + list(self.netD_multiresolution)

    def _resized(self, img, level):
        if level == 0:
            return img
        sh, sw = self.real_A.size(2) // (2 ** level), self.real_A.size(3) // (2 ** level)
        return F.ResizeNCHWFn.apply(img, sh, sw)

    def _real_A_pyramid(self):
        return [self._resized(self.real_A, i) for i in range(len(self._scales()))]

    def _gan(self, pyr_A, img_B, target_is_real, detach):
        """sum over scales of GANLoss(D_i(cat(A_i, B_i)))"""
        total = None
        for i, netD in enumerate(self._scales()):
            b = img_B.detach() if detach else img_B
            pred = netD.forward_engine(pyr_A[i], self._resized(b, i))
            term = self.criterionGAN.engine(pred, target_is_real)
            total = term if total is None else total + term
        return total

    def _gan_groups(self, pyr_A, imgs_B, targets_are_real, detach):
        """[sum over scales of GANLoss(D_i(cat(A_i, B_i)), target)  for B, target in zip(imgs_B, targets)], with ONE
        discriminator pass per scale over the batch-concatenation of all the (A, B) pairs"""
        total = None
        for i, netD in enumerate(self._scales()):
            imgs = []
            for b in imgs_B:
                imgs += [pyr_A[i], self._resized(b.detach() if detach else b, i)]
            pred = netD.forward_engine(*imgs, groups=len(imgs_B))
            terms = self.criterionGAN.engine_groups(pred, targets_are_real)
            total = terms if total is None else total + terms
        return [total[j] for j in range(len(imgs_B))]

    def _batch_d(self):
        return bool(getattr(self.opt, "batch_d", 1)) and self.opt.gan_mode == "lsgan"

    def backward_T_and_R(self):
        opt = self.opt
        pyr_A = self._real_A_pyramid()
        self.loss_L1_TR = F.L1Fn.apply(self.fake_TR_B, self.real_B, opt.lambda_recon).squeeze(0)
        self.loss_L1_RT = F.L1Fn.apply(self.fake_RT_B, self.real_B, opt.lambda_recon).squeeze(0)
        if self._batch_d():
            g_tr, g_rt = self._gan_groups(pyr_A, [self.fake_TR_B, self.fake_RT_B], [True, True], detach=False)
            self.loss_GAN_TR, self.loss_GAN_RT = opt.lambda_GAN * g_tr, opt.lambda_GAN * g_rt
        else:
            self.loss_GAN_TR = opt.lambda_GAN * self._gan(pyr_A, self.fake_TR_B, True, detach=False)
            self.loss_GAN_RT = opt.lambda_GAN * self._gan(pyr_A, self.fake_RT_B, True, detach=False)
        self.loss_smoothness = opt.lambda_smooth * self.stn_reg_term
        loss = self.loss_L1_TR + self.loss_L1_RT + self.loss_GAN_TR + self.loss_GAN_RT + self.loss_smoothness
        loss.backward()
        # the engine delivers parameter gradients inside its own backward calls (no AccumulateGrad node runs), so
        # autograd's end-of-backward stream sync does not cover the STN's stream: join it before the Adam launch
        self._join_stn_stream()
        return loss

    def backward_D(self):
        pyr_A = self._real_A_pyramid()
        if self._batch_d():
            loss_D_real, self.loss_D_fake_TR, self.loss_D_fake_RT = self._gan_groups(
                pyr_A, [self.real_B, self.fake_TR_B, self.fake_RT_B], [True, False, False], detach=True)
        else:
            loss_D_real = self._gan(pyr_A, self.real_B, True, detach=True)
            self.loss_D_fake_TR = self._gan(pyr_A, self.fake_TR_B, False, detach=True)
            self.loss_D_fake_RT = self._gan(pyr_A, self.fake_RT_B, False, detach=True)
        self.loss_D = 0.5 * self.opt.lambda_GAN * (loss_D_real + self.loss_D_fake_TR + self.loss_D_fake_RT)
        self.loss_D.backward()
        return self.loss_D

    def optimize_parameters(self):
        """One training iteration (reference nemar_model.py:266-288).  With --cuda_graph 1 the whole step (forward,
        both backward passes, both gradient all-reduces, both Adam launches: ~900 kernel launches) is captured once
        and replayed as a single graph launch."""
        if getattr(self.opt, "cuda_graph", 0) and self.device.type == "cuda":
            return self._optimize_parameters_graphed()
        return self._optimize_parameters_eager()

    def _optimize_parameters_graphed(self):
        st = self.__dict__.setdefault("_graph_state", {"eager_steps": 0, "graph": None, "failed": False})
        if st["failed"]:
            return self._optimize_parameters_eager()
        lrs = (self.optimizer_TR.param_groups[0]["lr"], self.optimizer_D.param_groups[0]["lr"])
        if st["graph"] is not None and st.get("lrs") != lrs:
            st["graph"] = None     # update_learning_rate() ran: the captured Adam launches carry the old rate -> re-capture
        if st["graph"] is not None:
            if self.real_A is not self._graph_inputs[0]:
                # a batch of another shape (e.g. the last, smaller batch of an epoch): set_input could not copy it into
                # the captured buffers, so this step is launched eagerly
                return self._optimize_parameters_eager()
            st["graph"].replay()
            L.COUNTERS["launches"] += st["launches"]     # the engine calls one replay stands for
            self.__dict__.update(st["outs"])             # (an eager step in between rebinds loss_* / image attributes)
            return
        if st["eager_steps"] >= 3 and self.real_A is not self._graph_inputs[0]:
            # the batch of the capture step has another shape (the short last batch of a small dataset): set_input bound
            # temporaries, and a graph captured now would read THEM forever.  Run it eagerly, capture on a later step.
            return self._optimize_parameters_eager()
        if st["eager_steps"] < 3:            # warm-up: lazy initialisation, allocator pools, packed-weight caches
            st["eager_steps"] += 1
            if self.__dict__.get("_graph_inputs") is None:
                self._graph_inputs = (self.real_A.clone(), self.real_B.clone())
                self.real_A, self.real_B = self._graph_inputs
            # warm up on a side stream: autograd binds every parameter's gradient accumulator to the stream that is
            # current when it is first used, and a capture must not depend on the legacy default stream
            side = st.setdefault("side", torch.cuda.Stream())
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._optimize_parameters_eager()
            torch.cuda.current_stream().wait_stream(side)
            return
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = L.COUNTERS["launches"]
            # capture on the SAME side stream the warm-up ran on: every autograd node (and every parameter's gradient
            # accumulator) is then bound to the capturing stream, so the backward pass needs no cross-stream event
            # (an event recorded outside the capture and waited on inside it is cudaErrorStreamCaptureIsolation)
            with torch.cuda.graph(graph, stream=st.get("side"),
                                  capture_error_mode=os.environ.get("NEMAR_GRAPH_CAPTURE_MODE", "thread_local")):
                self._optimize_parameters_eager()
            st["graph"] = graph
            st["lrs"] = lrs
            st["launches"] = L.COUNTERS["launches"] - n0
            # the tensors a replay writes: losses, images, the regulariser term (attributes bound during the capture)
            st["outs"] = {k: v for k, v in self.__dict__.items() if torch.is_tensor(v)}
            graph.replay()                    # the capture itself does not execute the step
        except Exception as e:               # noqa: BLE001 - any capture problem => eager launches
            print("CUDA graph capture failed (%s); continuing with eager launches" % str(e).splitlines()[0])
            st["failed"] = True
            torch.cuda.synchronize()
            return self._optimize_parameters_eager()

    def _optimize_parameters_eager(self):
        if self.device.type == "cuda":
            F.begin_step(self.device)          # device step counter: dropout masks change per step, also under replay
        self.forward()
        # D phase
        self.set_requires_grad([self.netT, self.netR], False)
        self.optimizer_D.zero_grad()
        self.backward_D()
        self.optimizer_D.step()           # all-reduce of the D bucket + flat Adam
        self.set_requires_grad([self.netT, self.netR], True)
        # T + R phase (uses the UPDATED discriminator, as the reference does)
        self.set_requires_grad([self.netD, *self.netD_multiresolution], False)
        self.optimizer_TR.zero_grad()
        self.backward_T_and_R()
        self.optimizer_TR.step()          # all-reduce of the T+R bucket + flat Adam
        self.set_requires_grad([self.netD, *self.netD_multiresolution], True)
        F.end_step()
