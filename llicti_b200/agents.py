"""Agent layer of the `main.py <config.json>` entry point for `mode: eval_model`
(mirrors agents/base.py:13-150 and agents/llicti_agent.py:14-164 of the reference: device
selection, checkpoint loading by the reference's file layout and key names, the per-image
compress -> rate table -> decompres -> lossless check loop and its log lines).

Training / validation modes are outside the B200 hot path and raise NotImplementedError.
"""
import logging
import shutil
import time

import torch

from .image_dl import TestImageLoader
from .model import LLICTI
from .rate import CompressionRLossList, RateLogger


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
            self.logger.info("Checkpoint loaded successfully from '{}' at (epoch {}) at (iteration {})".format(
                self.config.checkpoint_dir, self.current_epoch, self.current_iteration))
        except OSError:
            self.logger.info("!!! No checkpoint exists from '{}'. Continuing with available parameters...".format(
                self.config.checkpoint_dir))

    def save_checkpoint(self, filename="checkpoint.pth.tar", is_best=0):
        state = {"epoch": self.current_epoch, "iteration": self.current_iteration,
                 "best_valid_loss": self.best_valid_loss, "state_dict": self.model.state_dict()}
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
            elif mode in ("train", "debug", "test", "flops_est"):
                raise NotImplementedError(f"mode '{mode}' is outside the B200 compress/decompress path")
            else:
                raise NameError("'" + mode + "' is not a valid training mode.")
        except KeyboardInterrupt:
            self.logger.info("You have entered CTRL+C.. Wait to finalize")

    def finalize(self):
        self.logger.info("Please wait while finalizing the operation.. Thank you")
        self.save_checkpoint()


class LLICTIAgent(BaseAgent):
    def __init__(self, config):
        super().__init__(config)
        assert config.wtr_type == "lazydwt"
        self.model = LLICTI(config).to(self.device)
        self.compr_loss = CompressionRLossList()
        self.test_logger = RateLogger()
        self.test_loader = TestImageLoader(config.test_data)
        if config.mode in ("test", "validate", "debug", "eval_model"):
            self.load_checkpoint("model_best.pth.tar")
        self.model_size_estimation()

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
        for batch_idx, x in enumerate(self.test_loader):
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
        if self.test_logger.rate:
            self.test_logger.display(lr=0.0, typ="te")

    @torch.no_grad()
    def validate(self):
        """The reference's validate() (agents/llicti_agent.py:85-103) without its learning-rate scheduler: the rate
        estimate of forward() (`llicti_forward_dev`) over the validation crops, logged as the 'va' table."""
        import torch.nn.functional as F
        from .image_dl import ValidImageLoader
        from .rate import TrainRLossList
        self.model.eval()
        valid_data = getattr(self.config, "valid_data", None) or self.config.test_data      # (the shipped configs carry no training keys)
        loader = ValidImageLoader(valid_data, getattr(self.config, "val_patch_size", 0), getattr(self.config, "val_batch_size", 1))
        train_loss, valid_logger = TrainRLossList(), RateLogger()
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
        return valid_rate_loss + valid_rate2_loss

    def model_size_estimation(self, print_params=False):
        psz = sum(p.nelement() * p.element_size() for p in self.model.parameters())
        bsz = sum(b.nelement() * b.element_size() for b in self.model.buffers())
        if print_params:
            for name, p in self.model.named_parameters():
                print(name, tuple(p.size()))
        self.logger.info(" model param+buffer=total size: {:.3f}+{:.3f}={:.3f}MB".format(
            psz / 1024 ** 2, bsz / 1024 ** 2, (psz + bsz) / 1024 ** 2))
