import enum


class ParallelMode(enum.Enum):
    GLOBAL = 'global'
    DATA = 'data'
