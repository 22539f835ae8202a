"""Plugin loader (reference: kwatsch/get_trainer.py:23-85): settings.yaml names the network module / class and the
trainer module / class as strings; they are resolved with import_module + getattr.  Same signature, same returns
(``trainer`` when ``args_dict`` is given, ``(trainer, args_dict)`` when loading an experiment directory)."""
import os
from importlib import import_module

from kwatsch.common import load_settings


def get_trainer_dynamic(args_dict=None, src_path=None, model_nbr=None, eval_mode=False, args_only=False,
                        model_file=None, **kwargs):
    if model_nbr is None and args_dict is None:
        raise ValueError("ERROR - get_trainer - args_dict or model_filename needs to be specified")
    if model_file is not None:
        print("Warning - get trainer - RETRAIN model {}".format(model_file))
    model_file_sr = None
    if src_path is not None:
        src_path = os.path.expanduser(src_path)
        args_dict = load_settings(os.path.join(src_path, "settings.yaml"))
        if 'output_dir' not in args_dict.keys():
            args_dict['output_dir'] = src_path
        model_file = os.path.expanduser(os.path.join(src_path, "models", str(model_nbr) + ".models"))
        model_nbr_sr = kwargs.get("model_nbr_sr", None)
        if model_nbr_sr is not None:
            model_file_sr = os.path.expanduser(os.path.join(src_path, "models", str(model_nbr_sr) + ".models"))
    ae_class_name = "VanillaACAI" if 'ae_class' not in args_dict.keys() else \
        args_dict['ae_class'].replace('default', 'VanillaACAI')
    args_dict.setdefault('use_extra_latent_loss', False)
    args_dict.setdefault('use_alpha_probe', False)
    args_dict.setdefault('alpha_dims', None)
    if args_only:
        return None, args_dict
    ae_module = import_module(args_dict['module_network_path'].replace("/", ".").replace(".py", ""))
    ae_class = getattr(ae_module, ae_class_name)
    ae_model = ae_class(args_dict).to(args_dict['device'])
    aesr_model = None if model_file_sr is None else ae_class(args_dict).to(args_dict['device'])
    trainer_module = args_dict['module_trainer_path'].replace("/", ".").replace(".py", "")
    trainer_module = import_module(trainer_module.replace("utils.", "kwatsch."))    # legacy module name
    trainer_class = getattr(trainer_module, args_dict.get('trainer_class', "AEBaseTrainer"))
    trainer = trainer_class(args_dict, ae_model, model_file=model_file, eval_mode=eval_mode, model_sr=aesr_model,
                            model_file_sr=model_file_sr)
    if src_path is None:
        return trainer
    return trainer, args_dict
