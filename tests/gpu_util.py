"""Helpers for the -m gpu parity tests: build the CUDA trainers from oracle / golden weights."""
import numpy as np
import torch


class Box(object):
    def __init__(self, dim, low=-1.0, high=1.0):
        self.low = np.full((dim,), low, dtype=np.float32)
        self.high = np.full((dim,), high, dtype=np.float32)
        self.shape = (dim,)


def producers(O, A, H, q_out=1):
    from oac_explore_b200.networks import get_policy_producer, get_q_producer
    return get_policy_producer(O, A, [H, H]), get_q_producer(O, A, [H, H], output_size=q_out)


def load_net(net, sd):
    net.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})


def net_cpu(net):
    return {k: v.detach().to('cpu') for k, v in net.state_dict().items()}


def to_dev(batch):
    return {k: v.cuda() for k, v in batch.items()}
