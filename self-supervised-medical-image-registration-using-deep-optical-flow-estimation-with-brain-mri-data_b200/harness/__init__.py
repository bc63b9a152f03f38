"""Measurement harness around the hot path (host models, synthetic data, the training step)."""
from .pwcnet import PWCNetHost
from .registration import RegistrationStep, synthetic_pairs, make_optimizer, train_step, GraphedStep

__all__ = ["PWCNetHost", "RegistrationStep", "synthetic_pairs", "make_optimizer", "train_step", "GraphedStep"]
