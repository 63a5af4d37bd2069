"""iopath.common.file_io: local-filesystem PathManager (recsys/datasets/criteo.py:18,118)."""
import os


class PathManager:
    def open(self, path, mode='r', **kwargs):
        return open(path, mode)

    def exists(self, path):
        return os.path.exists(path)

    def ls(self, path):
        return os.listdir(path)


class PathManagerFactory:
    _managers = {}

    def get(self, key='') -> PathManager:
        if key not in self._managers:
            self._managers[key] = PathManager()
        return self._managers[key]
