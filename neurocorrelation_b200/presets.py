"""Headless restatement of the reference's presets (src/main.cpp:80-210) — the application logic
*above* the hot path: which inputs exist and how their rates are modulated between steps.

Every function takes a "brain" that exposes the driver verbs shared by the CUDA-backed host class
binding (neurocorrelation_b200.NeuCor), the reference harness and the CPU oracle
(`set_inputs`, `set_rate`, `set_params`, `enable_sweep`, `step`) and a `rand()` callable that
draws from the libc generator the reference uses (the preset code and the core share one
stream, main.cpp:45,102 and NeuCor.cpp:604-607).  All arithmetic is done in float32 exactly as
the C++ expressions are typed.
"""
import math

import numpy as np

F = np.float32
RAND_MAX = 2147483647
DT_DEFAULT = 0.0625  # 2**-4 ms: exactly representable, close to the GUI's 4 ms/s at 60 fps


def random_unit(rand):
    """static_cast<float>(rand()) / static_cast<float>(RAND_MAX)  (main.cpp:45, NeuCor.cpp:12-14)"""
    return F(rand()) / F(RAND_MAX)


def random_rate(rand):
    """SIMULATIONS::randomRate, main.cpp:44-46"""
    return random_unit(rand) * F(75.0)


def standard_inputs(rand):
    """Rates, positions and radii of the STANDARD preset, main.cpp:86-92 (three firers, radius 0.8,
    at radius 2 in the z=0 plane, 120 degrees apart; cosf/sinf evaluated in float32)."""
    rates = np.array([random_rate(rand), random_rate(rand), random_rate(rand)], F)
    radii = np.array([0.8, 0.8, 0.8], F)
    ang = [F(0.0), F(2.0944), F(4.1888)]
    pos = np.array([[F(math.cos(float(a))) * F(2.0), F(math.sin(float(a))) * F(2.0), F(0.0)] for a in ang], F)
    # (the same float32 positions are handed to the reference, the oracle and the CUDA engine)
    return rates, pos, radii


def standard_on_frame(rates, rand):
    """The STANDARD preset's per-frame input random walk, main.cpp:100-105 (3 rand() per call;
    inputs[1] is tied to inputs[0] — the "correlated" pair; inputs[2] is the uncorrelated one)."""
    for i in range(len(rates)):
        v = rates[i] + (random_unit(rand) - F(0.5)) * F(2.0)
        rates[i] = min(max(v, F(0.0)), F(75.0))
    rates[1] = rates[0]
    return rates


class StandardDriver:
    """Config C1 (BASELINE.json configs[0]): default main.cpp network run headless in sweep mode.

    Usage: build the brain after `srand(seed)`, then `drv = StandardDriver(brain, rand)`;
    `srand(777)` (separates construction and stepping streams); `drv.step()` x K.
    """

    def __init__(self, brain, rand, dt=DT_DEFAULT, learning_rate=1.0):
        self.brain = brain
        self.rand = rand
        self.rates, self.pos, self.radii = standard_inputs(rand)
        brain.set_inputs(self.rates, self.pos, self.radii)
        brain.enable_sweep()
        brain.set_params(dt, learning_rate, False)

    def step(self):
        standard_on_frame(self.rates, self.rand)
        for i, v in enumerate(self.rates):
            self.brain.set_rate(i, v)
        return self.brain.step()
