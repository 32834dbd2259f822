"""Rotating project.log, same entry point as the reference (simcore/logger_config.py:7-32)."""
import logging
import os
from logging.handlers import RotatingFileHandler


def get_logger(log_dir: str = ".", name: str = "dcsim_b200", level=logging.INFO):
    os.makedirs(log_dir, exist_ok=True)
    logger = logging.getLogger(f"{name}:{os.path.abspath(log_dir)}")
    if not logger.handlers:
        handler = RotatingFileHandler(os.path.join(log_dir, "project.log"), maxBytes=5_000_000, backupCount=2)
        handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(message)s"))
        logger.addHandler(handler)
        logger.setLevel(level)
        logger.propagate = False
    return logger
