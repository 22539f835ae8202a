"""Trainer classes -- drop-in for ``kwatsch/base_trainer.py``, ``kwatsch/trainer_ae.py``,
``kwatsch/cardiac/trainer_ae.py`` and ``kwatsch/brain/trainer_ae.py`` (hot-path subset).

Same constructor (``Trainer(args, ae, max_grad_norm=0, model_file=None, eval_mode=False, **kwargs)``), same methods
and attributes the entry scripts use (``train / validate / encode / decode / predict / save_models / load /
end_epoch_processing / reset_losses / save_losses / show_loss_on_tensorboard``, ``losses``, ``losses_test``,
``mean_losses*``, ``opt_ae``, ``percept_criterion``, ``epoch``, ``iters``), same checkpoint layout -- but ``train`` runs
the fused sm_100a step of ``training/engine.py`` instead of autograd, and reads all logged scalars back with one
device->host copy instead of five ``.item()`` syncs.
"""
from __future__ import annotations

import os
from collections import defaultdict

import numpy as np
import torch
from torch import optim

from . import ops, ops_train as T
from .lpips_b200 import PerceptualLoss
from .training.engine import TrainEngine


class BaseTrainer(object):
    # ------------------------------------------------------------------ init helpers (kwatsch/base_trainer.py:18-65)
    def _init_scheduler(self):
        if "use_lr_scheduler" in self.args.keys() and self.args['use_lr_scheduler']:
            self.opt_sched_ae = optim.lr_scheduler.CosineAnnealingLR(self.opt_ae, self.args["lr_iter_max"], eta_min=0,
                                                                     last_epoch=-1)

    def _init_percept_loss(self):
        self.ae_loss_func = "perceptual" if self.args.get('use_percept_loss', False) else "mse"
        if self.ae_loss_func == "perceptual":
            raise NotImplementedError("aesr_b200: use_percept_loss=True (LPIPS as reconstruction loss) is not an "
                                      "ae_combined configuration (networks/net_config.py:24)")
        need = not self.eval_model and (self.args.get('image_mix_loss_func') == "perceptual")
        if need:
            dev = self.args['device'] if str(self.args['device']).startswith('cuda') else 'cuda:0'
            self.percept_criterion = PerceptualLoss(model='net-lin', net='vgg', use_gpu=True, device=dev,
                                                    vgg_state=self.args.get('_vgg_state'),
                                                    vgg_weights=self.args.get('vgg_weights'),
                                                    random_init_seed=self.args.get('lpips_random_init_seed'))
        else:
            self.percept_criterion = None

    def _init_laploss(self):
        if not self.eval_model and self.args.get('use_laploss', False):
            raise NotImplementedError("aesr_b200: use_laploss is False in every ae_combined config "
                                      "(networks/net_config.py:25)")
        self.laploss = None

    def determine_image_mix_loss_func(self):
        if "image_mix_loss_func" in self.args.keys():
            self.image_mix_loss_func = self.args['image_mix_loss_func']
        else:
            self.image_mix_loss_func = "perceptual" if self.args.get('use_percept_loss') else "mse"

    # ------------------------------------------------------------------ eval-mode inference (base_trainer.py:216-323)
    def _use_sr_model(self, use_sr_model=False, **kwargs):
        if use_sr_model and self.model_sr is not None:
            return self.model_sr
        return self.model

    def _to_dev(self, x):
        return x.to(self._device) if not x.is_cuda else x

    def predict(self, x, eval=True, chunk_size=16, clear_cache=False, **kwargs):
        model = self.model
        model.eval() if eval else model.train()
        return model(self._to_dev(x))        # 180 GB of HBM: the reference's 256^2 CPU chunking is not needed

    def encode(self, x, eval=True, clear_cache=False, chunk_size=16, **kwargs):
        model = self._use_sr_model(kwargs.get("use_sr_model", False))
        model.eval() if eval else model.train()
        return model.encode(self._to_dev(x))

    def decode(self, z, eval=True, clear_cache=False, chunk_size=16, **kwargs):
        model = self._use_sr_model(kwargs.get("use_sr_model", False))
        model.eval() if eval else model.train()
        return model.decode(self._to_dev(z))

    # ------------------------------------------------------------------ losses
    def _mse_value(self, a, b) -> torch.Tensor:
        acc = torch.zeros(1, dtype=torch.float32, device=a.device)
        T.mse(a.float().contiguous(), b.float().contiguous(), acc)
        return acc

    def get_loss(self, reference, recons, is_test=False, store_loss=True):
        """F.mse_loss(recons, reference) (kwatsch/base_trainer.py:164-198, ae_loss_func == 'mse')."""
        loss = self._mse_value(self._to_dev(recons), self._to_dev(reference))[0]
        if store_loss:
            (self.losses_test if is_test else self.losses)['loss_ae_dist'].append(loss.item())
        return {"loss_ae": loss, 'loss_ae_dist': loss, "loss_laploss": 0}

    def _get_mixup_latent(self, **kwargs):
        z = kwargs.get('z')
        b = z.size(0) // 2
        w = torch.full((b,), 0.5, dtype=torch.float32, device=z.device)
        idx = torch.arange(b, dtype=torch.int32, device=z.device)
        _, z_mix = ops.lerp_latents(z.float().contiguous(), idx, idx + b, w, w, want_nchw=True)
        return z_mix

    def get_latent_loss(self, **kwargs):
        reference, z = kwargs.get('reference'), kwargs.get('z')
        z_mix = self._get_mixup_latent(**kwargs)
        z_ref = self.encode(reference, eval=True)
        return {'loss_latent': self._mse_value(z_mix, z_ref)[0], 'z_mix': z_mix}

    def validate(self, validation_batch, image_dict=None, frame_id=8, generate_images=True):
        """kwatsch/base_trainer.py:67-99 (image grids are visual logging, out of scope: returned as None)."""
        self.model.eval()
        img = self._to_dev(validation_batch['image'])
        z = self.encode(img, eval=True)
        img_recons = self.decode(z, eval=True)
        loss = self.get_loss(img, img_recons, is_test=True)['loss_ae']
        latent = self.get_latent_loss(reference=self._to_dev(validation_batch['slice_between']), z=z)
        self.test_predictions = {'z': z.detach().cpu(), 'img_recons': img_recons.detach().cpu(), 'z_device': z.device}
        self.losses_test['loss_ae'].append(loss.item())
        self.losses_test['loss_latent_1'].append(latent['loss_latent'].item())
        if self.epoch > self.args['epoch_threshold'] and 'vae' not in self.args['model']:
            self.save_best_val_model()
        return {"img_grid_recons": None, "loss_ae": self.losses_test['loss_ae'][-1]}

    def save_best_val_model(self, **kwargs):
        key = 'loss_ae_dist'
        if len(self.mean_losses_test[key]) > 1 and \
                np.argmin(self.mean_losses_test[key]) + 1 == len(self.mean_losses_test[key]):
            self.save_models(os.path.join(self.args['dir_models'], 'ae.models'), self.epoch + 1)

    # ------------------------------------------------------------------ checkpoints (base_trainer.py:353-362)
    def save_models(self, fname, epoch):
        torch.save({'model_dict_ae': self.model.state_dict(), 'optimizer_dict_ae': self.opt_ae.state_dict(),
                    'epoch': epoch}, fname)

    def load(self, fname):
        state_dict = torch.load(fname, map_location=self._device)
        self.model.load_state_dict(state_dict['model_dict_ae'])
        self.opt_ae.load_state_dict(state_dict['optimizer_dict_ae'])
        if self.engine is not None:
            self.engine.reload_optimizer_state()
        print("INFO - {} Loaded model parameters from {}".format(self.__class__.__name__, fname))

    # ------------------------------------------------------------------ bookkeeping (base_trainer.py:364-460)
    @property
    def iters(self):
        return self._iters

    def init_tensorboard(self, output_directory):
        from torch.utils.tensorboard.writer import SummaryWriter
        self.tb_writer = SummaryWriter(log_dir=os.path.join(output_directory, "tb"), comment=str(self.args))

    def show_loss_on_tensorboard(self, eval_type='train'):
        if eval_type == "train":
            loss_dict, mean_losses = self.losses, self.mean_losses
            self.loss_iters.append(self.iters)
        else:
            loss_dict, mean_losses = self.losses_test, self.mean_losses_test
        for loss_key in loss_dict.keys():
            mean_value = np.mean(np.array(loss_dict[loss_key]))
            if self.args['log_tensorboard']:
                self.tb_writer.add_scalar("{}/{}".format(loss_key, eval_type), mean_value, self.iters)
            mean_losses[loss_key].append(mean_value)

    def generate_train_images(self, **kwargs):
        pass                                  # PNG grids: visual logging, out of scope

    def end_epoch_processing(self, **kwargs):
        epoch = kwargs.get('epoch')
        if self.epoch > self.args['epoch_threshold']:
            self.save_models(os.path.join(self.args['dir_models'], '{:0d}.models'.format(epoch)), epoch)
        self.save_losses()
        self.epoch += 1

    def save_model(self, **kwargs):
        epoch, with_iters = kwargs.get('epoch'), kwargs.get('with_iters', False)
        name = '{:0d}.models'.format(epoch) if not with_iters else '{:0d}_{}.models'.format(epoch, self.iters)
        self.save_models(os.path.join(self.args['dir_models'], name), epoch)

    @staticmethod
    def load_losses(path_to_exper):
        path_to_exper = os.path.expanduser(path_to_exper)
        iters = np.load(os.path.join(path_to_exper, "loss_iters.npz"))['loss_iters']
        tr = np.load(os.path.join(path_to_exper, "losses_train.npz"))
        te = np.load(os.path.join(path_to_exper, "losses_test.npz"))
        return iters, {k: tr[k] for k in tr.files}, {k: te[k] for k in te.files}

    def save_losses(self):
        np.savez(os.path.join(self.args['output_dir'], "loss_iters.npz"), loss_iters=np.array(self.loss_iters))
        np.savez(os.path.join(self.args['output_dir'], "losses_train.npz"), **self.mean_losses)
        np.savez(os.path.join(self.args['output_dir'], "losses_test.npz"), **self.mean_losses_test)

    def init_weight_annealing(self, epochs):
        """kwatsch/base_trainer.py:456-459: reversed sigmoid ramp of ex_loss_weight1 (float64 like the reference)."""
        x = np.linspace(-5, 5, epochs)
        y = torch.sigmoid(torch.from_numpy(x)) * self.args.get('ex_loss_weight1', 0.001)
        self.loss_weights = y.numpy()[::-1]

    def reset_losses(self):
        for d in (self.losses, self.losses_test):
            for key in d.keys():
                d[key] = []


class AEBaseTrainer(BaseTrainer):
    """kwatsch/trainer_ae.py:16-109 -- plain ``ae`` step (MSE only) and the constructor every ae_combined trainer
    inherits."""
    combined = False

    def __init__(self, args, ae, max_grad_norm=0, model_file=None, eval_mode=False, **kwargs):
        super(AEBaseTrainer, self).__init__()
        self.args = args
        self.model = ae
        self.model_sr = kwargs.get('model_sr', None)
        self.eval_model = eval_mode
        self.model_file = model_file
        self.do_chunk = False
        self.eval_fixed_coeff = True
        self._device = next(ae.parameters()).device
        momentum = 0.9 if 'momentum' not in args.keys() else args['momentum']
        self.opt_ae = optim.Adam(self.model.parameters(), lr=args['lr'], weight_decay=args['weight_decay'],
                                 betas=(momentum, 0.999))
        self.opt_sched_ae = None
        self._init_scheduler()
        self.losses, self.losses_test = defaultdict(list), defaultdict(list)
        self.loss_iters = list()
        self.mean_losses, self.mean_losses_test = defaultdict(list), defaultdict(list)
        self.train_predictions, self.test_predictions = None, None
        self._iters = 1
        self.max_grad_norm = max_grad_norm
        if max_grad_norm:
            raise NotImplementedError("aesr_b200: max_grad_norm is 0 in every hot-path config (net_config.py:30)")
        self.use_multiple_gpu = False           # the reference's 2-GPU loss offload is replaced by data parallelism
        self.alpha05 = torch.tensor([0.5], dtype=torch.float32, device=self._device)[:, None, None, None]
        self._init_laploss()
        self._init_percept_loss()
        self.determine_image_mix_loss_func()
        self.ssim_criterion = None
        self.epoch = 0
        self.init_weight_annealing(self.args['epochs'])
        if self.args.get('use_ssim_loss'):
            raise NotImplementedError("ERROR - Disabled SSIM as loss when upgrading pytorch to 1.9 version!")
        self.engine = None if eval_mode else TrainEngine(self.model, self.opt_ae, sync_bn=bool(args.get('sync_bn', False)))
        if model_file is not None:
            self.load(model_file)
        if self.model_sr is not None and kwargs.get("model_file_sr", None) is not None:
            self.model_file_sr = kwargs.get("model_file_sr")
            self.load_caisr(self.model_file_sr)

    def load_caisr(self, fname):
        self.model_sr.load_state_dict(torch.load(fname, map_location=self._device)['model_dict_ae'])

    # -- per-trainer pieces -------------------------------------------------------------------------------------
    def _mix_weights(self, batch_item, B):
        w = torch.full((B,), 0.5, dtype=torch.float32, device=self._device)
        return w, w                              # alpha05 and (1 - alpha05), cardiac/trainer_ae.py:173

    def _extra_weight(self):
        if self.args.get('use_loss_annealing'):
            return float(self.loss_weights[self.epoch])
        return float(self.args.get('ex_loss_weight1', 0.0))

    def train(self, batch_item, keep_predictions=True, eval_mode=False):
        if eval_mode:
            raise NotImplementedError("aesr_b200: train(eval_mode=True) is not used by the entry scripts")
        if self.args.get('use_extra_latent_loss') or self.args.get('get_masks'):
            raise NotImplementedError("aesr_b200: use_extra_latent_loss / get_masks are off in every hot-path config")
        x = batch_item['image'].to(self._device, non_blocking=True)
        sb = batch_item['slice_between'].to(self._device, non_blocking=True)
        self.model.train()
        self._iters += 1
        B = x.shape[0] // 2
        wa, wb = self._mix_weights(batch_item, B)
        g = self.opt_ae.param_groups[0]
        res = self.engine.step(x, sb, wa, wb, lpips=self.percept_criterion, ex_loss_weight=self._extra_weight(),
                               combined=self.combined, lr=g['lr'], betas=g['betas'], eps=g['eps'],
                               weight_decay=g['weight_decay'], keep=keep_predictions)
        if self.opt_sched_ae is not None:
            self.opt_sched_ae.step()
        logs = TrainEngine.logged_losses(res)
        for k in ('loss_ae_dist', 'loss_ae_extra', 'loss_ae_dist_extra', 'loss_ae', 'loss_latent_1'):
            if k in logs:
                self.losses[k].append(logs[k])
        if not self.combined:
            self.model.eval()                    # the reference leaves the model in eval here
        if keep_predictions:
            mix = res['s_between_mix']
            if mix is None:                      # plain AE: no-grad decode of the 0.5/0.5 mix (trainer_ae.py:101-103)
                mix = self.decode(res['z_mix'], eval=not self.model.training)
            mix = mix.detach().cpu()
            self.train_predictions = {'z_mix': res['z_mix'].detach().cpu(), 'pred_alphas': torch.FloatTensor([0.5]),
                                      'slice_inbetween_mix': mix, 'slice_inbetween_05': mix,
                                      "reconstruction": res['reconstruction'].detach().cpu()}


class AETrainerEndToEnd(AEBaseTrainer):
    """kwatsch/cardiac/trainer_ae.py:8-182 (ACDC ae_combined): MSE + w * LPIPS(slice_between, dec(0.5 z1 + 0.5 z2))."""
    combined = True

    def validate(self, validation_batch, image_dict=None, frame_id=8, generate_images=True):
        val = super().validate(validation_batch, image_dict=image_dict, frame_id=frame_id,
                               generate_images=generate_images)
        z = self.test_predictions['z'].to(self._device)
        sb = self._to_dev(validation_batch['slice_between'])
        B = z.shape[0] // 2
        wa, wb = self._mix_weights(validation_batch, B)
        idx = torch.arange(B, dtype=torch.int32, device=self._device)
        self.model.eval()
        z16, z_mix = ops.lerp_latents(z.float().contiguous(), idx, idx + B, wa, wb, want_nchw=True)
        s_mix = self.model.decode_nhwc_eval(z16)
        z_ref = self.encode(sb, eval=True)
        self.losses_test['loss_latent_1'].append(self._mse_value(z_mix, z_ref).item())
        if self.percept_criterion is not None:
            extra = self._extra_weight() * self.percept_criterion(sb, s_mix, normalize=True).mean().item()
            self.losses_test['loss_ae_extra'].append(extra)
            self.losses_test['loss_ae_dist_extra'].append(extra)
        self.model.train()
        return val

    def save_best_val_model(self, **kwargs):
        key = 'loss_ae_dist_extra'
        if len(self.mean_losses_test[key]) > 1 and \
                np.argmin(self.mean_losses_test[key]) + 1 == len(self.mean_losses_test[key]):
            self.save_models(os.path.join(self.args['dir_models'], 'caisr.models'), self.epoch + 1)


class AETrainerBrain(AEBaseTrainer):
    """kwatsch/brain/trainer_ae.py:47-87 (plain ``ae`` on OASIS/dHCP)."""
    combined = False


class AETrainerExtension1Brain(AETrainerEndToEnd):
    """kwatsch/brain/trainer_ae.py:90-282 (OASIS / dHCP ae_combined): per-sample alpha_from / alpha_to [B,1]."""

    def _mix_weights(self, batch_item, B):
        wa = batch_item['alpha_from'].to(self._device).float().reshape(-1).contiguous()
        wb = batch_item['alpha_to'].to(self._device).float().reshape(-1).contiguous()
        return wa, wb
