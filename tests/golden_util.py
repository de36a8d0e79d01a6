"""Helpers shared by the parity tests: load a golden case frozen from the reference."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.dims4 = tuple(int(x) for x in z["dims4"])
        self.train, self.test = z["train"], z["test"]
        self.temps = z["temps"]
        self.S = int(z["S"])
        self.swap_interval = int(z["swap_interval"])
        self.do_swaps = bool(z["do_swaps"])
        self.swap_u = list(z["swap_u"])
        self.swap_log = [(int(a), int(b), bool(c)) for a, b, c in z["swap_log"]]
        self.n = len(self.temps)
        self.experiment_type = int(z["experiment_type"]) if "experiment_type" in z else 2

    def rnd(self, j):
        z = self.z
        return dict(w0_randn=z[f"c{j}_rnd_w0_randn"], lx=z[f"c{j}_rnd_lx"], noise=z[f"c{j}_rnd_noise"],
                    eta_noise=z[f"c{j}_rnd_eta_noise"], u=z[f"c{j}_rnd_u"])

    def theta_init(self, j):
        return self.z[f"c{j}_theta_init"]

    def get(self, j, key):
        return self.z[f"c{j}_{key}"]
