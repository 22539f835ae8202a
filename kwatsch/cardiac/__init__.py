"""Drop-in import paths of the reference trainers (settings.yaml: module_trainer_path, kwatsch/get_trainer.py:71-76)."""
