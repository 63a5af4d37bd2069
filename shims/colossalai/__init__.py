"""Stand-in for the parts of ColossalAI that /root/reference/recsys/dlrm_main.py:16-24,377-378 and
recsys/models/dlrm.py:14-23 use: the launcher, the default argument parser, the rank-filtered logger, the global
context, and -- the product -- the cached embedding-bag layers (colossalai.nn.parallel.layers -> cachedembedding_b200)."""
import argparse
import os
import random

import numpy as np
import torch
import torch.distributed as dist

from . import logging  # noqa: F401
from . import core, context, nn  # noqa: F401

__version__ = "0.0.0+cachedembedding_b200.shim"


def get_default_parser():
    """The launcher flags every ColossalAI script accepts (upstream colossalai.get_default_parser)."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--config', type=str, help='path to the config file')
    parser.add_argument('--host', type=str, help='the master address for distributed training')
    parser.add_argument('--port', type=int, help='the master port for distributed training')
    parser.add_argument('--world_size', type=int, help='world size for distributed training')
    parser.add_argument('--rank', type=int, help='rank for the default process group')
    parser.add_argument('--local_rank', type=int, help='local rank on the node')
    parser.add_argument('--backend', type=str, default='nccl', help='backend for distributed communication')
    return parser


def launch_from_torch(config=None, backend: str = 'nccl', seed: int = 1024, verbose: bool = True):
    """One process per GPU, started by torchrun: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* come from the environment."""
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world_size = int(os.environ.get('WORLD_SIZE', 1))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    else:
        backend = 'gloo'
    if not dist.is_initialized():
        kwargs = {}
        if backend == 'nccl':
            kwargs['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world_size, **kwargs)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    core.global_context._set(rank, local_rank, world_size)
    if verbose and rank == 0:
        logging.get_dist_logger().info(f'Distributed environment is initialized, world size: {world_size}')
