"""colossalai.logging: a logger whose methods take `ranks=[...]` (recsys/dlrm_main.py:20,332-333,390-391)."""
import logging as _logging

import torch.distributed as dist


class DistributedLogger:
    def __init__(self, name='colossalai'):
        self._logger = _logging.getLogger(name)
        if not self._logger.handlers:
            handler = _logging.StreamHandler()
            handler.setFormatter(_logging.Formatter('[%(asctime)s] %(levelname)s %(message)s'))
            self._logger.addHandler(handler)
        self._logger.setLevel(_logging.INFO)
        self._logger.propagate = False

    @staticmethod
    def _mine(ranks):
        if ranks is None or not (dist.is_available() and dist.is_initialized()):
            return True
        return dist.get_rank() in ranks

    def info(self, message, ranks=None):
        if self._mine(ranks):
            self._logger.info(message)

    def warning(self, message, ranks=None):
        if self._mine(ranks):
            self._logger.warning(message)

    def error(self, message, ranks=None):
        if self._mine(ranks):
            self._logger.error(message)

    def debug(self, message, ranks=None):
        if self._mine(ranks):
            self._logger.debug(message)


_LOGGER = None


def get_dist_logger(name='colossalai'):
    global _LOGGER
    if _LOGGER is None:
        _LOGGER = DistributedLogger(name)
    return _LOGGER


def disable_existing_loggers(include=None, exclude=('colossalai',)):
    for name in list(_logging.root.manager.loggerDict):
        if not any(name.startswith(e) for e in exclude):
            _logging.getLogger(name).setLevel(_logging.WARNING)
