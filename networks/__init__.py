"""Drop-in import path: settings.yaml names the model module as ``networks/acai_vanilla.py`` (kwatsch/get_trainer.py:61-68)."""
