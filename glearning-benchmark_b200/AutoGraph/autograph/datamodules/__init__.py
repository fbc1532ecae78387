"""Shim namespace so `from autograph.datamodules.data.tokenizer import Graph2TrailTokenizer` resolves (trainer/train_agtt.py:17-21)."""
