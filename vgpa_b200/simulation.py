"""
Simulation: the reference's orchestration (src/var_bayes/simulation.py:23-347) over
the CUDA path -- same `setup(params, data)` / `run()` flow, the same sim_params JSON
schema (vgpa_main.py:38-40) and the same output keys.  `main(params_file)` is the
entry point of vgpa_main.py.  Results are saved as a compressed .npz (h5py is not
in this image); the key set is the reference's HDF5 dataset set.
"""
import json
import time
from pathlib import Path

import numpy as np

from .dynamics import dynamical_systems
from .likelihood import GaussianLikelihood
from .ode import BwdOde, FwdOde
from .prior import PriorKL0
from .scg import SCG
from .variational import VarGP

REQUIRED = ("Output_Name", "Model", "Ode-method", "Time-window", "Noise", "Observations", "Drift",
            "Prior", "Random-Seed")


def validate_input_parameters_file(filename):
    """vgpa_main.py:14-59"""
    filename = Path(filename)
    if not filename.is_file():
        raise ValueError(f" File {filename} doesn't exist.")
    with open(filename, "r") as fh:
        params = json.load(fh)
    for k in REQUIRED:
        if k not in params:
            raise ValueError(f" Key: {k}, is not given.")
    return params


class Simulation(object):

    def __init__(self, name=None, device=0):
        self.name = str(name) if name else "ID_None"
        self.m_data = {}
        self.output = {}
        self.device = device
        self.scg_stats = None

    def setup(self, params, data=None):
        """simulation.py:92-178"""
        md = self.m_data
        md["drift"], md["noise"] = params["Drift"], params["Noise"]
        md["time_window"], md["ode_solver"] = params["Time-window"], params["Ode-method"]
        md["random_seed"], md["obs_setup"] = params["Random-Seed"], params["Observations"]
        md["mu0"], md["tau0"] = params["Prior"]["mu0"], params["Prior"]["tau0"]
        key = str(params["Model"]).upper()
        if key not in dynamical_systems:
            raise ValueError(f" Simulation: Unknown stochastic model -> {key}")
        model = dynamical_systems[key](md["noise"]["sys"], md["drift"]["theta"], md["random_seed"])
        md["model"], md["single_dim"] = model, model.single_dim
        tw = md["time_window"]
        model.make_trajectory(tw["t0"], tw["tf"], tw["dt"])
        if data is not None:
            md["obs_t"], md["obs_y"], md["obs_noise"] = data[0], data[1], md["noise"]["obs"]
        else:
            md["obs_t"], md["obs_y"], md["obs_noise"] = model.collect_obs(
                md["obs_setup"]["density"], md["noise"]["obs"], md["obs_setup"]["operator"])
        if md["single_dim"]:
            md["m0"] = model.sample_path[0] + 0.1 * model.rng.standard_normal()
            md["s0"] = 0.2
        else:
            dim_d = model.sample_path.shape[-1]
            md["m0"] = model.sample_path[0] + 0.1 * model.rng.standard_normal(dim_d)
            md["s0"] = 0.2 * np.eye(dim_d)
            md["mu0"] = md["mu0"] * np.ones(dim_d)
            md["tau0"] = md["tau0"] * np.eye(dim_d)

    def build(self):
        """The constructor block of simulation.py:189-212."""
        md = self.m_data
        dt = md["time_window"]["dt"]
        fwd = FwdOde(dt, md["ode_solver"], md["single_dim"], self.device)
        bwd = BwdOde(dt, md["ode_solver"], md["single_dim"], self.device)
        lik = GaussianLikelihood(md["obs_y"], md["obs_t"], md["obs_noise"], md["obs_setup"]["operator"],
                                 md["single_dim"], self.device)
        kl0 = PriorKL0(md["mu0"], md["tau0"], md["single_dim"])
        return VarGP(md["model"], md["m0"], md["s0"], fwd, bwd, lik, kl0, md["obs_y"], md["obs_t"],
                     device=self.device)

    def run(self, max_it=500, display=True, optimizer="host"):
        """simulation.py:180-267.  optimizer="host": the SCG loop on the host, as in the reference (13 MB
        numpy vectors per operation at the Lorenz-96 shape, x and grad F crossing PCIe at every evaluation);
        optimizer="device": the same optimiser with x, the gradients and the search direction resident in
        HBM (vgpa_b200.batched_scg.BatchedSCG with a batch of one) -- the same trace to 1e-6
        (tests/test_gpu_scg.py), without the host vector arithmetic and the transfers."""
        vgpa = self.build()
        options = {"max_it": max_it, "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": display}
        x0 = vgpa.initialization()
        t0 = time.perf_counter()
        if optimizer == "device":
            from .batched_scg import BatchedSCG
            optimize = BatchedSCG(vgpa._ev, options)
            X, fxs = optimize(x0)
            x, fx = X[0].cpu().numpy(), float(fxs[0])
            n = int(optimize.stats["MaxIt"][0])
            self.scg_stats = {"MaxIt": n, "fx": optimize.stats["fx"][:, 0].copy(), "dfx": optimize.stats["dfx"][:, 0].copy(),
                              "beta": optimize.stats["beta"][:, 0].copy(), "f_eval": float(optimize.stats["f_eval"][0]),
                              "df_eval": float(optimize.stats["df_eval"][0])}
            del optimize, X
        elif optimizer == "host":
            optimize = SCG(vgpa.free_energy, vgpa.gradient, options)
            x, fx = optimize(x0.copy())
            self.scg_stats = optimize.stats
        else:
            raise ValueError(f" Simulation.run: unknown optimizer {optimizer!r} (host, device).")
        print(f" Elapsed time: {(time.perf_counter() - t0):.2f} seconds.")
        md = self.m_data
        if md["model"].single_dim:
            n = md["model"].sample_path.size
            self.output["at"], self.output["bt"] = x[:n], x[n:]
        else:
            n, d = md["model"].sample_path.shape
            self.output["at"] = x[:n * d * d].reshape(n, d, d)
            self.output["bt"] = x[n * d * d:].reshape(n, d)
        self.output["fx"] = fx
        vgpa.free_energy(x)            # make the cached state the one of the returned x
        self.output.update(vgpa.arg_out)
        vgpa.close()

    def save(self):
        """simulation.py:269-311: every entry of `output` as one gzip-compressed dataset of
        `<name>.h5` (scalars as 1-D arrays).  Without h5py (it is not a dependency of the CUDA
        path) the same keys go to `<name>.npz`; `load` reads either."""
        if not self.output:
            print(f" {self.__class__.__name__}: Simulation data structure 'output' is empty.")
            return
        stem = self.name.strip().replace(" ", "_")
        data = {k: np.atleast_1d(v) if np.isscalar(v) else np.asarray(v) for k, v in self.output.items()}
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            out = Path(stem + ".h5")
            with h5py.File(out, "w") as out_file:
                for key, val in data.items():
                    out_file.create_dataset(key, data=val, shape=val.shape, compression="gzip")
        else:
            out = Path(stem + ".npz")
            np.savez_compressed(out, **data)
        print(f" Saved the results to: {out}")


def load(filename=None):
    """simulation.py:316-345: the dictionary written by Simulation.save (.h5 or .npz)."""
    if filename is None:
        raise RuntimeError(" load_data: No input file is given.")
    path = Path(filename)
    if path.suffix == ".npz":
        with np.load(path, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}
    import h5py
    with h5py.File(path, "r") as input_file:
        return {key: np.array(input_file[key]) for key in input_file}


def main(params_file=None, data_file=None):
    """vgpa_main.py:62-145"""
    import sys
    if params_file is None:
        print(" The simulation can't run without input parameters.")
        sys.exit(1)
    try:
        params = validate_input_parameters_file(params_file)
    except ValueError as e0:
        print(e0)
        sys.exit(1)
    obs_data = None
    if data_file is not None:
        arr = np.loadtxt(data_file, delimiter=",")
        obs_data = (arr[:, 0].astype(int), arr[:, 1:].squeeze())
    try:
        sim = Simulation(params["Output_Name"] or "Sim_00")
        sim.setup(params, obs_data)
        sim.run()
        sim.save()
    except Exception as e1:
        print(e1)
        sys.exit(1)
