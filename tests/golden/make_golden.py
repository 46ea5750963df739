"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array stored here is an input fed to, or an output produced by, the
reference's own classes (imported through oracle/ref_import.py with the gym /
matplotlib / gtimer shims and the torch-1.4 "Mode A" optimizer patch, see that
file).  The oracle (oracle/oac_oracle.py) and the CUDA path are both tested
against these files; /root/reference is never read at test time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import as ri  # noqa: E402
from tests.util import synth_batch, synth_eps  # noqa: E402


def digest(x):
    """Compact fingerprint of a tensor: (sum, sum of squares, first, middle, last) in fp64."""
    x = np.asarray(x.detach() if isinstance(x, torch.Tensor) else x, dtype=np.float64).ravel()
    return np.array([x.sum(), (x * x).sum(), x[0], x[x.size // 2], x[-1]])


def net_arrays(prefix, net, out, full):
    for k, v in net.state_dict().items():
        out['%s/%s' % (prefix, k)] = v.detach().numpy().copy() if full else digest(v)


def save(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **d)
    print("wrote %s (%d arrays, %.1f KB)" % (name, len(d), os.path.getsize(path) / 1024.))


def sac_case(name, O, A, B, hidden, n_steps, full, seed=0):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, hidden=hidden)
    torch.manual_seed(seed)
    tr = ref.trainer.SACTrainer(pp, qp, action_space=ac_space, policy_lr=3e-4, qf_lr=3e-4,
                                soft_target_tau=5e-3, use_automatic_entropy_tuning=True)
    ri.mode_a(tr)
    out = dict(meta=np.array([O, A, B, n_steps, seed] + list(hidden)))
    nets = dict(policy=tr.policy, qf1=tr.qf1, qf2=tr.qf2, target_qf1=tr.target_qf1,
                target_qf2=tr.target_qf2)
    for n, net in nets.items():
        net_arrays('init/' + n, net, out, full)
    diag = []
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        with ri.injected_noise(eps):
            tr._need_to_update_eval_statistics = True
            tr.train_from_torch(dict(batch))
        es = tr.eval_statistics
        diag.append([es['QF1 Loss'], es['QF2 Loss'], es['Policy Loss'], es['Alpha'],
                     es['Log Pis Mean'], es['Q Targets Mean'], es['Q1 Predictions Mean'],
                     es['Q2 Predictions Mean'], es['Policy mu Mean'], es['Policy log std Mean']])
    out['diag'] = np.array(diag, dtype=np.float64)
    for n, net in nets.items():
        net_arrays('final/' + n, net, out, full)
    out['final/log_alpha'] = tr.log_alpha.detach().numpy().copy()
    save(name, out)


def poac_case(name, O, A, B, P, hidden, n_steps, share_layers, counts, seed=1):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, hidden=hidden, q_out=P if share_layers else 1)
    torch.manual_seed(seed)
    tr = ref.particle_trainer_oac.ParticleTrainer(
        pp, qp, n_estimators=P, action_space=ac_space, share_layers=share_layers, counts=counts,
        deterministic=False, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
        use_automatic_entropy_tuning=True, delta=0.95, q_min=0.0, q_max=500.0)
    ri.mode_a(tr)
    out = dict(meta=np.array([O, A, B, n_steps, seed, P, int(share_layers), int(counts)] + list(hidden)))
    net_arrays('init/policy', tr.policy, out, True)
    for i in range(len(tr.qfs)):
        net_arrays('init/qf%d' % i, tr.qfs[i], out, True)
        net_arrays('init/tf%d' % i, tr.tfs[i], out, True)
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=20 + s, counts=counts)
        eps = synth_eps(2, B, A, seed=200 + s)   # eps[0] -> next_obs draw, eps[1] -> obs draw
        with ri.injected_noise(eps):
            tr.train_from_torch(dict(batch))
    net_arrays('final/policy', tr.policy, out, True)
    for i in range(len(tr.qfs)):
        net_arrays('final/qf%d' % i, tr.qfs[i], out, True)
        net_arrays('final/tf%d' % i, tr.tfs[i], out, True)
    out['final/log_alpha'] = tr.log_alpha.detach().numpy().copy()
    save(name, out)


def goac_case(name, O, A, B, hidden, n_steps, share_layers, counts, seed=2):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, hidden=hidden, q_out=2 if share_layers else 1)
    torch.manual_seed(seed)
    tr = ref.gaussian_trainer.GaussianTrainer(
        pp, qp, n_estimators=2, action_space=ac_space, share_layers=share_layers, counts=counts,
        policy_lr=3e-4, qf_lr=3e-4, std_lr=3e-5, soft_target_tau=5e-3, delta=0.95, q_min=0.0,
        q_max=500.0)
    out = dict(meta=np.array([O, A, B, n_steps, seed, int(share_layers), int(counts)] + list(hidden)))
    nets = dict(policy=tr.policy, target_policy=tr.target_policy, q=tr.q, q_target=tr.q_target)
    if not share_layers:
        nets.update(std=tr.std, std_target=tr.std_target)
    for n, net in nets.items():
        net_arrays('init/' + n, net, out, True)
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=30 + s, counts=counts)
        tr.train_from_torch(dict(batch))
    for n, net in nets.items():
        net_arrays('final/' + n, net, out, True)
    save(name, out)


def explore_case(name, O, A, hidden, n_obs, seed=4):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, hidden=hidden)
    torch.manual_seed(seed)
    tr = ref.trainer.SACTrainer(pp, qp, action_space=ac_space)
    hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
    out = dict(meta=np.array([O, A, n_obs, seed] + list(hidden)), beta_UB=np.array(4.66),
               delta=np.array(23.53))
    for n, net in dict(policy=tr.policy, qf1=tr.qf1, qf2=tr.qf2).items():
        net_arrays('init/' + n, net, out, True)
    rng = np.random.RandomState(0)
    obs, eps2, acts, mus = [], [], [], []
    for i in range(n_obs):
        ob = rng.randn(O)
        torch.manual_seed(50 + i)
        ac, _ = ref.optimistic_exploration.get_optimistic_exploration_action(
            ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp)
        torch.manual_seed(50 + i)
        torch.normal(torch.zeros(A), torch.ones(A))          # policy(ob) rsample, discarded
        eps2.append(torch.normal(torch.zeros(A), torch.ones(A)).numpy())
        mu, _ = ref.optimistic_exploration.get_optimistic_exploration_action(
            ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp, deterministic=True)
        obs.append(ob), acts.append(ac), mus.append(mu)
    out.update(obs=np.array(obs), eps_sample=np.array(eps2), action=np.array(acts),
               mu_E_deterministic=np.array(mus))
    save(name, out)


def explore_ensemble_case(name, O, A, P, hidden, n_obs, seed=5):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, hidden=hidden, q_out=P)
    torch.manual_seed(seed)
    tr = ref.particle_trainer_oac.ParticleTrainer(pp, qp, n_estimators=P, action_space=ac_space,
                                                  share_layers=True, q_min=0., q_max=500.,
                                                  deterministic=False)
    hp = dict(beta_UB=4.66, delta=20.53, share_layers=True)
    out = dict(meta=np.array([O, A, n_obs, seed, P] + list(hidden)), beta_UB=np.array(4.66),
               delta=np.array(20.53))
    net_arrays('init/policy', tr.policy, out, True)
    net_arrays('init/qf0', tr.qfs[0], out, True)
    rng = np.random.RandomState(1)
    obs, eps2, acts = [], [], []
    for i in range(n_obs):
        ob = rng.randn(O)
        torch.manual_seed(60 + i)
        ac, _ = ref.optimistic_exploration.get_optimistic_exploration_action(
            ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp)
        torch.manual_seed(60 + i)
        torch.normal(torch.zeros(A), torch.ones(A))
        eps2.append(torch.normal(torch.zeros(A), torch.ones(A)).numpy())
        obs.append(ob), acts.append(ac)
    out.update(obs=np.array(obs), eps_sample=np.array(eps2), action=np.array(acts))
    save(name, out)


def replay_case(name, O=7, A=3, N=50, T=130, B=16):
    ref = ri.load_reference()
    ob_space, ac_space = ri.make_spaces(O, A)
    rb = ref.replay_buffer.ReplayBufferCount(N, ob_space, ac_space)
    rng = np.random.RandomState(0)
    samples = dict(obs=[], act=[], rew=[], nobs=[], term=[])
    out = dict(meta=np.array([O, A, N, T, B]))
    nb = 0
    for t in range(T):
        o, a, r, no, d = rng.randn(O), rng.rand(A), rng.randn(), rng.randn(O), bool(rng.rand() < 0.1)
        rb.add_sample(observation=o, action=a, reward=r, next_observation=no, terminal=d, env_info={})
        for k, v in zip(('obs', 'act', 'rew', 'nobs', 'term'), (o, a, r, no, d)):
            samples[k].append(v)
        if t % 17 == 5:
            np.random.seed(t)
            idx = np.random.randint(0, rb._size, B)
            np.random.seed(t)
            b = rb.random_batch(B)
            out['batch%d/t' % nb] = np.array(t)
            out['batch%d/indices' % nb] = idx
            for k, v in b.items():
                out['batch%d/%s' % (nb, k)] = v
            nb += 1
    for k, v in samples.items():
        out['stream/' + k] = np.array(v)
    out['n_batches'] = np.array(nb)
    out['final/top_size'] = np.array([rb._top, rb._size])
    out['final/counts'] = rb._counts.copy()
    save(name, out)


if __name__ == "__main__":
    assert ri.reference_available(), "needs /root/reference"
    sac_case("sac_small.npz", 11, 3, 32, (32, 32), 3, full=True)
    sac_case("sac_riverswim.npz", 1, 1, 256, (256, 256), 3, full=False)
    sac_case("sac_humanoid.npz", 376, 17, 256, (256, 256), 3, full=False)
    poac_case("poac_shared_small.npz", 11, 3, 32, 5, (32, 32), 3, True, False)
    poac_case("poac_shared_counts_small.npz", 11, 3, 32, 5, (32, 32), 3, True, True)
    poac_case("poac_separate_small.npz", 11, 3, 32, 4, (32, 32), 3, False, False)
    goac_case("goac_shared_small.npz", 11, 3, 32, (32, 32), 3, True, False)
    goac_case("goac_shared_counts_small.npz", 11, 3, 32, (32, 32), 3, True, True)
    goac_case("goac_separate_small.npz", 11, 3, 32, (32, 32), 3, False, False)
    explore_case("explore_twin_small.npz", 11, 3, (32, 32), 6)
    explore_ensemble_case("explore_ensemble_small.npz", 11, 3, 5, (32, 32), 4)
    replay_case("replay_counts.npz")
