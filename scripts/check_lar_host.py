import ctypes, numpy as np, sys
sys.path.insert(0, ".")
from oracle import vmp
lib = ctypes.CDLL("tests/c/_lar_host.so")
def run(y, p, iters, params):
    T, B = y.shape
    y = np.ascontiguousarray(y, np.float32)
    ns = p + p * (p + 1) // 2
    ws = np.zeros((T, ns, B), np.float32)
    xm = np.zeros((T, p, B), np.float32); xc = np.zeros((T, p, p, B), np.float32)
    tm = np.zeros((iters, p, B), np.float32); tc = np.zeros((iters, p, p, B), np.float32)
    gs = np.zeros((iters, B), np.float32); gr = np.zeros((iters, B), np.float32)
    fe = np.zeros((iters, B), np.float64); st = np.zeros(B, np.int32)
    prm = np.asarray(params, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.lar_host_run(p, T, ctypes.c_longlong(B), iters, P(prm), P(y), P(ws), P(xm), P(xc), P(tm), P(tc), P(gs), P(gr), P(fe), P(st))
    assert rc == 0
    return dict(x_mean=xm, x_cov=xc, theta_mean=tm, theta_cov=tc, gamma_shape=gs, gamma_rate=gr, free_energy=fe, status=st)
st, obs = vmp.latent_ar_reference_data()
rng = np.random.default_rng(0)
Y = np.stack([obs, obs[::-1], obs + 0.1 * rng.standard_normal(500)], 1)
for p in (1, 2, 3, 5, 6):
    ref = vmp.latent_ar(Y, p, 5.0, 15)
    r = run(Y, p, 15, [5.0, 1, 1, 1, 1, 1, 1, 1])
    print(p, "status", r["status"], "fe err", np.abs(r["free_energy"] - ref["free_energy"]).max(), "last", r["free_energy"][-1],
          "x_mean", np.linalg.norm(r["x_mean"] - ref["x_mean"]) / np.linalg.norm(ref["x_mean"]),
          "x_cov", np.linalg.norm(r["x_cov"] - ref["x_cov"]) / np.linalg.norm(ref["x_cov"]),
          "theta", np.abs(r["theta_mean"] - ref["theta_mean"]).max(), "tcov", np.abs(r["theta_cov"] - ref["theta_cov"]).max(),
          "g", np.abs(r["gamma_shape"] / r["gamma_rate"] - ref["gamma_shape"] / ref["gamma_rate"]).max())
