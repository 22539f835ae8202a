"""Per-(model, dataset) architecture defaults merged into the CLI args (reference: networks/net_config.py:19-92).
Only the ``ae`` / ``ae_combined`` rows of the hot path are provided; other model families are out of scope."""

MODULE_PATH = {"VanillaACAI": "networks/acai_vanilla.py"}


class NetworkConfig(object):
    def __init__(self, network, dataset=None, ae_class="VanillaACAI"):
        self.network, self.dataset, self.ae_class = network, dataset, ae_class
        self.architecture = {}
        self.load_config()

    def load_config(self):
        a = self.architecture
        a.update(width=128, latent_width=16, depth=32, colors=1, latent=16, use_laploss=False, use_percept_loss=False,
                 n_res_block=None, use_batchnorm=True, use_sigmoid=True, max_grad_norm=0, fine_tune=False,
                 ex_loss_weight1=0.5)
        if self.ae_class not in MODULE_PATH:
            raise ValueError("Error - NetworkConfig - aesr_b200 provides ae_class VanillaACAI only, got {}".format(
                self.ae_class))
        a['module_network_path'] = MODULE_PATH[self.ae_class]
        brain = self.dataset in ['dHCP', 'ADNI', 'OASIS']
        if self.network in ("ae", "aesr"):
            if self.dataset is None or self.dataset == "ACDC":
                a['module_trainer_path'], a['trainer_class'] = "kwatsch/trainer_ae.py", "AEBaseTrainer"
            elif brain:
                a['module_trainer_path'], a['trainer_class'] = "kwatsch/brain/trainer_ae.py", "AETrainerBrain"
            else:
                raise ValueError("Error - NetworkConfig - Unsupported combination {}/{}".format(self.network, self.dataset))
            a['image_mix_loss_func'] = None
        elif self.network in ("ae_combined", "aesr_combined"):
            a['image_mix_loss_func'] = "perceptual"
            if self.dataset == "ACDC":
                a['module_trainer_path'], a['trainer_class'] = "kwatsch/cardiac/trainer_ae.py", "AETrainerEndToEnd"
            elif brain:
                a['module_trainer_path'], a['trainer_class'] = "kwatsch/brain/trainer_ae.py", "AETrainerExtension1Brain"
            else:
                raise ValueError("Error - NetworkConfig - Unsupported combination {}/{}".format(self.network, self.dataset))
        else:
            raise ValueError("Error - NetworkConfig - model family {} is outside the aesr_b200 hot path".format(self.network))
