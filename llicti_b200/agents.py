"""Agent layer of the `main.py <config.json>` entry point (mirrors agents/base.py:13-150 and
agents/llicti_agent.py:14-164 of the reference: device selection, checkpoint loading and saving by the reference's
file layout and key names, `mode: eval_model` -- the per-image compress -> rate table -> decompres -> lossless check
loop and its log lines --, `mode: validate`, `mode: train` / `debug`, `model_size`, `flops_est`, `test`).

`mode: train` is the reference's loop (llicti_agent.py:48-83, base.py:132-146) around the library's training step:
`self.model(x)` and `.backward()` run in libllicti_b200 (`llicti_forward_dev` / `llicti_backward_dev`, fp32); Adam,
gradient clipping, ReduceLROnPlateau and the checkpoint are torch's, as in the reference.
"""
import logging
import shutil
import time

import torch

from .image_dl import TestImageLoader
from .model import LLICTI
from .rate import CompressionRLossList, RateLogger
from .shard import average_gradients


class BaseAgent:
    def __init__(self, config):
        self.config = config
        self.logger = logging.getLogger("Agent")
        self.best_valid_loss = float("inf")
        self.current_epoch = 0
        self.current_iteration = 0
        self.manual_seed = config.seed
        if not (torch.cuda.is_available() and config.cuda):
            raise RuntimeError("the B200 path needs `cuda: true` and a CUDA device; there is no CPU fallback")
        self.cuda = True
        # one process per GPU under torchrun (data-parallel training): the rank's own device, NCCL for the gradient average
        import os
        self.world, self.rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
        if self.world > 1:
            config.gpu_device = int(os.environ.get("LOCAL_RANK", "0"))
            if not torch.distributed.is_initialized():
                torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(config.gpu_device)))
        self.device = torch.device("cuda", int(config.gpu_device))
        torch.cuda.set_device(self.device)
        torch.cuda.manual_seed(self.manual_seed)

    def load_checkpoint(self, filename):
        """experiments/<exp_name>/checkpoints/<filename>, dict with 'state_dict' (base.py:51-81).
        A missing file is tolerated like in the reference ("Continuing with available parameters")."""
        path = self.config.checkpoint_dir + filename
        try:
            self.logger.info("Loading checkpoint '{}'".format(path))
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            self.current_epoch = ckpt.get("epoch", 0)
            self.current_iteration = ckpt.get("iteration", 0)
            self.best_valid_loss = ckpt.get("best_valid_loss", self.best_valid_loss)
            self.model.load_state_dict(ckpt["state_dict"])
            if getattr(self.config, "resume_training", False) and self.config.mode == "train":          # base.py:61-75
                for key in ("optimizer", "scheduler", "train_logger", "trnit_logger", "valid_logger"):
                    if key in ckpt and getattr(self, key, None) is not None:
                        getattr(self, key).load_state_dict(ckpt[key])
            self.logger.info("Checkpoint loaded successfully from '{}' at (epoch {}) at (iteration {})".format(
                self.config.checkpoint_dir, self.current_epoch, self.current_iteration))
        except OSError:
            self.logger.info("!!! No checkpoint exists from '{}'. Continuing with available parameters...".format(
                self.config.checkpoint_dir))

    def save_checkpoint(self, filename="checkpoint.pth.tar", is_best=0):
        if getattr(self, "rank", 0) != 0:            # data-parallel training: the ranks hold the same weights, rank 0 writes
            return
        state = {"epoch": self.current_epoch, "iteration": self.current_iteration,
                 "best_valid_loss": self.best_valid_loss, "state_dict": self.model.state_dict()}
        for key in ("optimizer", "scheduler", "train_logger", "trnit_logger", "valid_logger"):   # base.py:88-92 (training runs)
            obj = getattr(self, key, None)
            if obj is not None:
                state[key] = obj.state_dict()
        torch.save(state, self.config.checkpoint_dir + filename)
        if is_best:
            shutil.copyfile(self.config.checkpoint_dir + filename, self.config.checkpoint_dir + "model_best.pth.tar")

    def run(self):
        mode = self.config.mode
        try:
            if mode == "eval_model":
                self.eval_model()
            elif mode == "model_size":
                self.model_size_estimation(print_params=True)
            elif mode == "validate":
                self.validate()
            elif mode in ("train", "debug"):      # debug: the reference wraps train() in autograd's anomaly detection (base.py:113-115),
                self.train()                      # which has nothing to inspect in a backward pass that is one library call
            elif mode == "test":
                self.test()
            elif mode == "flops_est":
                self.flops_estimation()
            else:
                raise NameError("'" + mode + "' is not a valid training mode.")
        except KeyboardInterrupt:
            self.logger.info("You have entered CTRL+C.. Wait to finalize")

    def train(self):
        """base.py:132-146."""
        for epoch in range(self.current_epoch, self.config.max_epoch):
            self.current_epoch = epoch
            self.train_one_epoch()
            if not (self.current_epoch + 1) % getattr(self.config, "validate_every", 1):
                valid_loss = self.validate()
                if valid_loss is not None:
                    is_best = valid_loss < self.best_valid_loss
                    if is_best:
                        self.best_valid_loss = valid_loss
                    self.save_checkpoint(is_best=is_best)
            self.current_epoch += 1

    def finalize(self):
        self.logger.info("Please wait while finalizing the operation.. Thank you")
        self.save_checkpoint()
        import os
        dump = os.environ.get("LLICTI_TEST_DUMP_WEIGHTS")          # tests: every rank's final weights ("{rank}" in the path)
        if dump:
            torch.save({k: v.detach().cpu() for k, v in self.model.state_dict().items()}, dump.format(rank=getattr(self, "rank", 0)))
        if getattr(self, "world", 1) > 1 and torch.distributed.is_initialized():
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()


class LLICTIAgent(BaseAgent):
    def __init__(self, config):
        super().__init__(config)
        assert config.wtr_type == "lazydwt"
        self.model = LLICTI(config).to(self.device)
        self.compr_loss = CompressionRLossList()
        self.test_logger = RateLogger()
        self.test_loader = TestImageLoader(config.test_data)
        self.optimizer = self.scheduler = self.train_logger = self.trnit_logger = self.valid_logger = None
        if config.mode == "train":                                   # llicti_agent.py:19-37
            from .image_dl import TrainImageLoader
            from .rate import TrainRLossList
            dirs = [getattr(config, "train_data_%d" % i) for i in range(1, int(getattr(config, "num_train_dirs", 1)) + 1)]
            self.train_loader = TrainImageLoader(dirs, config.patch_size, config.batch_size,
                                                 getattr(config, "patches_per_img", 1), seed=config.seed, rank=self.rank, world=self.world)
            self.train_loss = TrainRLossList()
            self.train_logger, self.trnit_logger, self.valid_logger = RateLogger(), RateLogger(), RateLogger()
            self.lr = config.learning_rate
            self.optimizer = torch.optim.Adam([{"params": self.model.parameters(), "lr": self.lr}])
            self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, factor=0.5, patience=16, threshold=0.0001,
                                                                        threshold_mode="rel", cooldown=15, min_lr=2.5e-05, eps=1e-08)
            self.grad_acc_iters = int(getattr(config, "grad_acc_iters", 1))
        if config.mode in ("test", "validate", "debug", "eval_model"):
            self.load_checkpoint("model_best.pth.tar")
        elif getattr(config, "resume_training", False):
            self.load_checkpoint(getattr(config, "checkpoint_file", "checkpoint.pth.tar"))
        if self.world > 1 and config.mode == "train":                # every rank starts from rank 0's weights (fresh initialisations differ)
            with torch.no_grad():
                for p in self.model.parameters():
                    torch.distributed.broadcast(p.data, 0)
        self.model_size_estimation()

    def train_one_epoch(self):
        """llicti_agent.py:48-83, line for line; `self.model(x)` and `.backward()` are the library's training step."""
        self.model.train()
        for batch_idx, x in enumerate(self.train_loader):
            x = x.to(self.device)
            if x.dim() == 5:
                x = x.view(-1, x.shape[2], x.shape[3], x.shape[4])
            self_infos_y_list = self.model(x)
            r_loss, rate1_list = self.train_loss.forward(torch.numel(x), self_infos_y_list)
            (r_loss / self.grad_acc_iters).backward()
            if (self.current_iteration + 1) % self.grad_acc_iters == 0:
                average_gradients(list(self.model.parameters()))      # identity on one GPU; the mean over the ranks under torchrun
                torch.nn.utils.clip_grad_value_(self.model.parameters(), clip_value=5.0)
                self.optimizer.step()
                self.optimizer.zero_grad()
            self.current_iteration += 1
            self.train_logger(rate1_list)
            self.trnit_logger(rate1_list)
            if not (self.current_iteration + 1) % getattr(self.config, "loss_prnt_iters", 2000):
                self.trnit_logger.display(lr=self.optimizer.param_groups[0]["lr"], typ="it")
                valid_loss = self.validate()
                self.model.train()
                if valid_loss is not None:
                    is_best = valid_loss < self.best_valid_loss
                    if is_best:
                        self.best_valid_loss = valid_loss
                    self.save_checkpoint(is_best=is_best)
        if self.train_logger.rate:
            self.train_logger.display(lr=self.optimizer.param_groups[0]["lr"], typ="tr")

    @torch.no_grad()
    def eval_model(self):
        self.model.eval()
        codec = self.model._codec()
        cc = self.model.codec_config
        self.logger.info(" B200 path: cnn_impl={} ({}), sub_len={} ({}), stream fingerprint {}".format(
            cc.cnn_impl, "tcgen05 tensor cores, {} operands".format("fp16" if codec.cnn_operands == 2 else "bf16") if cc.cnn_impl == 1 else "fp32 CUDA cores",
            cc.sub_len, "interleaved substreams" if cc.sub_len > 0 else "torchac-compatible streams",
            codec.fingerprint.hex()))
        codec.profile(True)          # per-kernel-class launch groups of this run, logged below
        first = 0
        if self.world > 1:           # one process per GPU (torchrun): a contiguous block of the test images per rank, no exchange
            from .shard import shard_range
            files = self.test_loader.files
            first, stop = shard_range(len(files), self.rank, self.world)
            self.test_loader.files = files[first:stop]
        for batch_idx, x in enumerate(self.test_loader):
            batch_idx += first
            x = x.to(self.device)
            text = "{:3d} {:3d}x{:3d} ".format(batch_idx, x.shape[2], x.shape[3])
            t0 = time.time()
            bytestream_list, _ = self.model.compress(x)
            enc_time = time.time() - t0
            self.test_logger(self.compr_loss.forward(torch.numel(x), bytestream_list))
            total = sum(len(b) * 8 for row in bytestream_list for b in row)
            t0 = time.time()
            x_reco = self.model.decompres(bytestream_list, self.device)
            dec_time = time.time() - t0
            maxerr = ((x - x_reco) * 255).abs().max().item()
            if maxerr >= 0.5:
                self.logger.info(text + "bpsp= {:.3f} Enc/Dec-Times:{:.3f}/{:.3f} (Error: Decoded img does NOT match "
                                        "original image perfectly! The maximum of absolute error is {:.4f})".format(
                                            total / torch.numel(x), enc_time, dec_time, maxerr))
            else:
                self.logger.info(text + "bpsp= {:.3f} Enc/Dec-Times:{:.3f}/{:.3f} "
                                        "(Check: Decoded img matches original)".format(total / torch.numel(x), enc_time,
                                                                                      dec_time))
        prof = codec.profile_read()
        codec.profile(False)
        self.logger.info(" B200 kernel classes launched: " + ", ".join(
            "{}{}={}".format(k, "[tcgen05]" if k == "cnn" and cc.cnn_impl == 1 else "", int(v[1])) for k, v in prof.items() if v[1]))
        if self.world > 1:           # the table is the mean over ALL images: sums and counts reduced over the ranks, rank 0 prints
            import numpy as np
            from .shard import reduce_stats
            rows = self.model.num_scales + 1
            local = np.asarray(self.test_logger.rate, dtype=np.float64).reshape(-1, rows, 9)
            sums, _ = reduce_stats(local.sum(axis=0).reshape(-1).tolist() + [float(local.shape[0])], [0.0], device=self.device)
            n_all = sums[-1]
            self.test_logger.rate = [np.asarray(sums[:-1]).reshape(rows, 9) / n_all] if (n_all and self.rank == 0) else []
        if self.test_logger.rate:
            self.test_logger.display(lr=0.0, typ="te")

    @torch.no_grad()
    def validate(self):
        """The reference's validate() (agents/llicti_agent.py:85-103): the rate estimate of forward()
        (`llicti_forward_dev`) over the validation crops, logged as the 'va' table; during training the result steps the
        learning-rate scheduler (:101)."""
        import torch.nn.functional as F
        from .image_dl import ValidImageLoader
        from .rate import TrainRLossList
        self.model.eval()
        valid_data = getattr(self.config, "valid_data", None) or self.config.test_data      # (the shipped configs carry no training keys)
        loader = ValidImageLoader(valid_data, getattr(self.config, "val_patch_size", 0), getattr(self.config, "val_batch_size", 1))
        train_loss, valid_logger = TrainRLossList(), (self.valid_logger or RateLogger())
        B = 2 ** (max(self.config.dwtlevels) + 1)
        for x in loader:
            x = x.to(self.device)
            h, w = x.size(2), x.size(3)
            x = F.pad(x, (0, (w + B - 1) // B * B - w, 0, (h + B - 1) // B * B - h), mode="replicate")   # _pad_img (:105-113)
            _, rate1_list = train_loss.forward(torch.numel(x), self.model(x))
            valid_logger(rate1_list)
        if not valid_logger.rate:
            self.logger.info(" validate: no images in {}".format(valid_data))
            return None
        valid_rate_loss, valid_rate2_loss = valid_logger.display(lr=0.0, typ="va")
        if self.scheduler is not None:
            self.scheduler.step(valid_rate_loss + valid_rate2_loss)
        return valid_rate_loss + valid_rate2_loss

    @torch.no_grad()
    def test(self):
        """The reference's test() is an empty stub (agents/llicti_agent.py:114-119, "test should be modified to have actual
        entropy coding"): eval_model is the mode that codes."""
        self.model.eval()

    def flops_estimation(self):
        """The reference asks ptflops for the multiply-accumulates of one forward() of a 3 x 512 x 512 image
        (agents/llicti_agent.py:194-200).  The count is closed-form here: per position of a scale, band b runs four
        sub-networks of (taps_b x chs + chs x chs + chs x 15) MACs, taps = 3 channels x kernel positions of the band's
        branches (LLICTI_nets.py:650-675), and scale s has (512 / 2^(s+1))^2 positions."""
        G, S = self.model.codec_config.chs, self.model.num_scales
        taps = (3 * 16, 3 * (12 + 12), 3 * (12 + 12 + 16))
        per_pos = [4 * (t * G + G * G + G * 15) for t in taps]
        positions = sum((512 >> (s + 1)) ** 2 for s in range(S))
        macs = positions * sum(per_pos)
        params = sum(p.nelement() for p in self.model.parameters())
        self.logger.info("{:<30}  {:.3f} GMac  (per position: {} + {} + {} MACs for the three bands; {} positions over {} scales)".format(
            "Computational complexity: ", macs / 1e9, *per_pos, positions, S))
        self.logger.info("{:<30}  {:.3f} M".format("Number of parameters: ", params / 1e6))
        return macs, params

    def model_size_estimation(self, print_params=False):
        psz = sum(p.nelement() * p.element_size() for p in self.model.parameters())
        bsz = sum(b.nelement() * b.element_size() for b in self.model.buffers())
        if print_params:
            for name, p in self.model.named_parameters():
                print(name, tuple(p.size()))
        self.logger.info(" model param+buffer=total size: {:.3f}+{:.3f}={:.3f}MB".format(
            psz / 1024 ** 2, bsz / 1024 ** 2, (psz + bsz) / 1024 ** 2))
