"""CPU fp64 restatement of the reference's Gaussian message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``rxinfer.jl_b200/``)
may import this; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker or as the timed CPU baseline.

The arithmetic of the reference hot path lives in un-vendored Julia packages
(ReactiveMP ~6.0.0, ExponentialFamily 2.1.0, BayesBase 1.5.0, FastCholesky 1.3.0;
/root/reference/Project.toml:43-73), so this oracle restates the published
algorithm of those rules and anchors on the reference's own call sites, tests and
fixed-data golden values (see ``tests/golden/`` and ``tests/test_oracle_goldens.py``).

Pin status:
  * Gamma-precision VMP + Addition + Normal rules: PINNED against
    test/models/aliases/aliases_gamma_tests.jl:43-44 (posterior mean bit-exact,
    Bethe free energy to 1e-12).
  * Two-node Gaussian BP + log-evidence: PINNED against
    test/models/models_tests.jl:242-336 (1.5 / 3.51551, 1.0 / 2.26551).
  * Normal entropy: PINNED against test/score/diagnostics_tests.jl:24.
  * LGSSM schedule: PINNED against the reference's own regression pins -- the test data of
    test/models/statespace/mlgssm_test.jl:70-97 (BFE 6275.9015944677) and ulgssm_tests.jl:27-32
    (BFE 1854.297647) are regenerated with oracle/julia_rng.py (StableRNGs.jl + Julia's ziggurat
    randn) and reproduced to 4e-9 / 7e-7; also cross-checked against textbook Kalman + RTS (1e-11).
  * HGF / GCV: PINNED -- the data stream of test/models/statespace/hgf_tests.jl:94-102 is regenerated
    (oracle/julia_rng.py) and the reference test runs verbatim: coverage / variance assertions (:121-133) and the
    free-energy pin 1.009879989585 +- 0.01 (:118; the oracle gives 1.0098705, see oracle/hgf.py).
  * Latent autoregressive model (AR node, structured VMP; oracle/vmp.py::latent_ar): PINNED -- the data of
    test/models/autoregressive/lar_tests.jl:128-157 is regenerated (StableRNG(123)) and the reference's free-energy pins are
    reproduced: Univariate AR(1) 518.918234267 vs 518.9182342 (:170, every printed digit), Multivariate AR(5) 514.65389 vs
    514.66086 +- 0.01 (:201, inside the reference's tolerance; residual 0.007 unexplained); all other assertions verbatim.
  * Autoregressive regression model (ar_tests.jl): the reference's assertions on its regenerated data (no absolute pin there).
  * Streaming mean-field Gamma model (test/inference/inference_tests.jl:752-860): rules pinned as above; the test's
    own assertion (free energy non-increasing over the iterations, :846) holds for oracle/vmp.py::stream_vmp_gamma.
"""
