"""``kwatsch.common`` settings IO (reference: kwatsch/common.py:45-68): settings.yaml is a YAML dump of the merged args."""
import argparse

import yaml


def load_settings(fname):
    with open(fname, 'r') as fp:
        return yaml.load(fp, Loader=yaml.FullLoader)


def save_settings(args, fname):
    with open(fname, 'w') as fp:
        yaml.dump(vars(args), fp)


def loadExperimentSettings(fname):
    with open(fname, 'r') as fp:
        return argparse.Namespace(**yaml.load(fp, Loader=yaml.FullLoader))


def saveExperimentSettings(args, fname):
    with open(fname, 'w') as fp:
        yaml.dump(args if isinstance(args, dict) else vars(args), fp)
