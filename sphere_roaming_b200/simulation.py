"""The reference's operator interface for this path, by name (src/simulation.h:65-104),
as batch calls over the C ABI.  Argument order and meaning follow the reference:

    read_config(filename)                                         simulation.h:65
    initialize_particles(he_number, conf, force_list)             simulation.h:76-79
    simulate_particles(conf, he_number, force_list, particles)    simulation.h:94-97
    save_particles(path, particles)                               simulation.h:104
    free_particles(particles)                                     simulation.h:85

``particles`` is a ``Batch``: what the reference holds as ``Particles[conf.number]`` with
one malloc'ed ``Particle[]`` each (main.c:55, simulation.c:310), flattened.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import engine, host
from .types import Config


@dataclass
class Batch:
    counts: np.ndarray          # particles per trajectory (Particles.no)
    particle: np.ndarray        # flat Particle records, trajectory after trajectory
    first_index: int = 0        # Particles.index of the first trajectory
    results: np.ndarray | None = None


_engines: dict = {}


def _engine_for(conf: Config, force_list: np.ndarray, mode: str) -> engine.Engine:
    key = (bytes(conf), force_list.ctypes.data, len(force_list), mode)
    if key not in _engines:
        _engines.clear()
        _engines[key] = engine.Engine(conf, force_list, mode=mode)
    return _engines[key]


def read_config(filename: str) -> Config:
    return host.read_config(filename)


def initialize_particles(he_number: int, conf: Config, force_list: np.ndarray, first_index: int = 0,
                         mode: str = "exact") -> Batch:
    """Sampling stage for ``conf.number`` trajectories (Philox stream keyed by seed, droplet
    size and global trajectory index)."""
    counts, flat = _engine_for(conf, force_list, mode).sample(he_number, conf.number, first_index=first_index)
    return Batch(counts, flat, first_index)


def simulate_particles(conf: Config, he_number: int, force_list: np.ndarray, particles: Batch,
                       mode: str = "exact") -> Batch:
    """The driver loop over ``simulate_particles`` (main.c:76-78), in place like the reference."""
    final, res = _engine_for(conf, force_list, mode).simulate(he_number, particles.counts, particles.particle)
    particles.particle = final
    particles.results = res
    return particles


def save_particles(path: str, particles: Batch) -> None:
    host.write_particlefile(path, particles.counts, particles.particle, particles.first_index)


def free_particles(particles: Batch) -> None:
    particles.particle = particles.particle[:0]
    particles.counts = particles.counts[:0]
