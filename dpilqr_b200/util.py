"""Host-side helpers of the drop-in API (layout conventions and scenario generation).

Joint vectors are agent-major concatenations with a uniform per-agent stride taken from
agent 0 (reference util.py:90-117, 229-236).  Scenario generation stays host NumPy so that a
seeded run consumes the global RNG streams exactly like the reference does
(reference util.py:125-226); it is measurement input, not part of the GPU hot path.
"""

import itertools
import random as _random
from dataclasses import dataclass
from pathlib import Path

import numpy as np

π = np.pi
repopath = Path(__file__).parent.parent.resolve()


@dataclass
class Point:
    """3-D point whose z defaults to 0 (reference util.py:20-45)."""

    x: float
    y: float
    z: float = 0

    @property
    def ndim(self):
        return 2 if self.z == 0 else 3

    def __add__(self, o):
        return Point(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Point(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, o):
        return Point(self.x * o.x, self.y * o.y, self.z * o.z)

    def __repr__(self):
        return str((self.x, self.y, self.z))

    def hypot2(self):
        return self.x**2 + self.y**2 + self.z**2


def _pair_indices(n_agents):
    return np.array(list(itertools.combinations(range(n_agents), 2)))


def compute_pairwise_distance(X, x_dims, n_d=2):
    """Distance between every pair of agents over the first ``n_d`` coordinates, rows = time
    (reference util.py:48-61).  Host utility for analysis; the solver's own distance
    evaluations run inside the CUDA kernels."""
    assert len(set(x_dims)) == 1
    n_agents, n_states = len(x_dims), x_dims[0]
    if n_agents == 1:
        raise ValueError("Can't compute pairwise distance for one agent.")
    pairs = _pair_indices(n_agents)
    Xa = np.asarray(X).reshape(-1, n_agents, n_states)
    diff = Xa[:, pairs[:, 0], :n_d] - Xa[:, pairs[:, 1], :n_d]
    return np.sqrt(np.add.reduce(diff * diff, axis=2))


def compute_pairwise_distance_nd(X, x_dims, n_dims, dec_ind=None):
    """Per-pair distance over min(n_dims[i], n_dims[j]) coordinates (reference util.py:64-87)."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(1, -1)
    n_states, n_agents = x_dims[0], len(x_dims)
    pairs = list(itertools.combinations(range(n_agents), 2))
    if dec_ind is not None:
        pairs = [p for p in pairs if dec_ind in p]
    out = np.zeros((X.shape[0], len(pairs)))
    for k, (i, j) in enumerate(pairs):
        nd = min(n_dims[i], n_dims[j])
        out[:, k] = np.linalg.norm(X[:, i * n_states:i * n_states + nd] - X[:, j * n_states:j * n_states + nd], axis=1)
    return out


def split_agents(Z, z_dims):
    """Partition joint states/controls into per-agent column blocks (reference util.py:90-92)."""
    return np.split(np.atleast_2d(Z), np.cumsum(z_dims[:-1]), axis=1)


def split_agents_gen(z, z_dims):
    """Generator over per-agent slices with the stride of agent 0 (reference util.py:95-99)."""
    dim = z_dims[0]
    for i in range(len(z_dims)):
        yield z[i * dim:(i + 1) * dim]


def split_graph(Z, z_dims, graph):
    """Column blocks of Z grouped per graph entry, in the order of each id list (reference util.py:102-117)."""
    assert len(set(z_dims)) == 1
    position = {id_: i for i, id_ in enumerate(list(graph))}
    width = z_dims[0]
    return [
        np.concatenate([Z[:, position[id_] * width:(position[id_] + 1) * width] for id_ in members], axis=1)
        for members in graph.values()
    ]


def pos_mask(x_dims, n_d=2):
    """Boolean mask of the position entries of a joint state (reference util.py:120-122)."""
    return np.array([i % x_dims[0] < n_d for i in range(sum(x_dims))])


def randomize_locs(n_pts, random=False, rel_dist=3.0, var=3.0, n_d=2):
    """Uniform random points, optionally pushed apart until every pair is farther than
    ``rel_dist`` (reference util.py:125-149).  Consumes np.random exactly like the reference."""
    push = 0.1 * n_pts
    x = var * np.random.uniform(-1, 1, (n_pts, n_d))
    if random:
        return x
    pairs = _pair_indices(n_pts)
    movers = np.arange(n_pts)
    while movers.size:
        center = np.mean(x, axis=0)
        dists = compute_pairwise_distance(x.flatten(), [n_d] * n_pts).T
        movers = pairs[dists.flatten() <= rel_dist]
        x[movers] += push * (x[movers] - center)
    return x


def face_goal(x0, xf):
    """Point the last state (heading) of every agent at its goal, with a little noise (reference util.py:152-162)."""
    noise = 0.01
    delta = xf[:, :2] - x0[:, :2]
    headings = np.arctan2(*np.rot90(delta, 1))
    x0[:, -1] = headings + noise * np.random.randn(x0.shape[0])
    xf[:, -1] = headings + noise * np.random.randn(x0.shape[0])
    return x0, xf


def random_setup(n_agents, n_states, is_rotation=False, n_d=2, energy=None, do_face=False, **kwargs):
    """Random initial / goal joint states as column vectors (reference util.py:165-195)."""
    x_i = randomize_locs(n_agents, n_d=n_d, **kwargs)
    if is_rotation:
        from scipy.spatial.transform import Rotation

        θ = π + _random.uniform(-π / 4, π / 4)
        R = Rotation.from_euler("z", θ).as_matrix()[:2, :2]
        x_f = x_i @ R - x_i.mean(axis=0)
    else:
        x_f = randomize_locs(n_agents, n_d=n_d, **kwargs)
    x0 = np.c_[x_i, np.zeros((n_agents, n_states - n_d))]
    xf = np.c_[x_f, np.zeros((n_agents, n_states - n_d))]
    if do_face:
        x0, xf = face_goal(x0, xf)
    x0, xf = x0.reshape(-1, 1), xf.reshape(-1, 1)
    if energy:
        x0 = normalize_energy(x0, [n_states] * n_agents, energy, n_d)
        xf = normalize_energy(xf, [n_states] * n_agents, energy, n_d)
    return x0, xf


def compute_energy(x, x_dims, n_d=2):
    """Sum of the agents' distances from the origin (reference util.py:198-200)."""
    return np.linalg.norm(x[pos_mask(x_dims, n_d)].reshape(-1, n_d), axis=1).sum()


def normalize_energy(x, x_dims, energy=10.0, n_d=2):
    """Centre the positions and scale them so compute_energy == energy (reference util.py:203-217)."""
    x = x.copy()
    mask = pos_mask(x_dims, n_d)
    center = x[mask].reshape(-1, n_d).mean(0)
    x[mask] -= np.tile(center, len(x_dims)).reshape(-1, 1)
    x[mask] *= energy / compute_energy(x, x_dims, n_d)
    assert x.size == sum(x_dims)
    return x


def perturb_state(x, x_dims, n_d=2, var=0.5):
    """Gaussian noise on the positions (reference util.py:220-226)."""
    x = x.copy()
    mask = pos_mask(x_dims, n_d)
    x[mask] += var * np.random.randn(*x[mask].shape)
    return x


def uniform_block_diag(*arrs):
    """Dense block-diagonal matrix of equally shaped blocks (reference util.py:229-236)."""
    rows, cols = arrs[0].shape
    out = np.zeros((len(arrs) * rows, len(arrs) * cols))
    for i, arr in enumerate(arrs):
        out[rows * i:rows * (i + 1), cols * i:cols * (i + 1)] = arr
    return out


def distance_to_goal(x, x_goal, n_agents, n_states, n_d):
    """Per-agent distance to the goal over the first n_d coordinates (reference util.py:239-240)."""
    return np.linalg.norm((x - x_goal).reshape(n_agents, n_states)[:, :n_d], axis=1)
