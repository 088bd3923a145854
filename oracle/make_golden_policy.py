"""TEST INFRASTRUCTURE (not product code).  Extracts the reference's shipped MLP controller and its published
evaluation into a plain-array fixture: tests/golden/mlp_controller.npz.

Sources (read-only, this container only; the GPU box has no /root/reference):
  gym_fixed_wing/examples/models/mlp_controller/model.pkl   stable-baselines 2 PPO2 save = a zip holding `parameters`
                                                            (np.savez of the tf variables) + `data` (json)
  .../mlp_controller/obs_rms.pkl, ret_rms.pkl               VecNormalize running statistics (pickled RunningMeanStd;
                                                            stable_baselines is not installed, so the class is stubbed)
  gym_fixed_wing/examples/evaluations/eval_res_RL_MLP_none.npy   the result dictionary evaluate_controller.py:169 saved
                                                            for this controller on test_set_wind_none (README table)

Run:  python -m oracle.make_golden_policy
"""
import io
import json
import os
import pickle
import zipfile

import numpy as np

REF = "/root/reference/gym_fixed_wing/examples"
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Stub:
    def __setstate__(self, st):
        self.__dict__.update(st)


class _Unpickler(pickle.Unpickler):
    def find_class(self, mod, name):
        if mod.startswith("stable_baselines"):
            return _Stub
        return super().find_class(mod, name)


def main():
    mdir = os.path.join(REF, "models", "mlp_controller")
    z = zipfile.ZipFile(os.path.join(mdir, "model.pkl"))
    par = np.load(io.BytesIO(z.read("parameters")))
    names = json.loads(z.read("parameter_list"))
    out = {}
    for k in names:
        # "model/pi_fc0/w:0" -> "pi_fc0_w"
        out[k.split(":")[0].replace("model/", "").replace("/", "_")] = np.asarray(par[k])
    with open(os.path.join(mdir, "obs_rms.pkl"), "rb") as f:
        o = _Unpickler(f).load()
    with open(os.path.join(mdir, "ret_rms.pkl"), "rb") as f:
        r = _Unpickler(f).load()
    out["obs_mean"], out["obs_var"] = np.asarray(o.mean, dtype=np.float64), np.asarray(o.var, dtype=np.float64)
    out["ret_var"] = np.float64(r.var)
    res = np.load(os.path.join(REF, "evaluations", "eval_res_RL_MLP_none.npy"), allow_pickle=True).item()
    out["pub_lengths"] = np.array([len(x) for x in res["rewards"]])
    out["pub_rewards"] = np.concatenate([np.asarray(x, dtype=np.float64) for x in res["rewards"]])
    for metric, per_state in res.items():
        if metric == "rewards":
            continue
        for state, vals in per_state.items():
            out["pub_%s_%s" % (metric, state)] = np.asarray(vals, dtype=np.float64)
    path = os.path.join(GOLDEN, "mlp_controller.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, sorted(out))


if __name__ == "__main__":
    main()
