#!/usr/bin/env python
"""Golden vectors for the particle-filter path: a LITERAL, scalar Python restatement of the
reference's Scala, written independently of oracle/ and of the package's arithmetic.

The reference cannot run here (Scala; no JVM in this image) and its own tests pin no number of this
path (src/test/scala/SamplingTest.scala:12-22 only checks lengths), so these vectors are a SECOND
restatement, not output of the reference: plain Python floats (IEEE double, like Scala's Double),
`math.exp/log/lgamma`, sequential left folds, and the ECDF as an actual sorted map with
"duplicate key overwrites" semantics (scala.collection.immutable.TreeMap).  They pin the C++
oracle against an implementation that shares no code with it.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.json

Citations: model/X.scala = src/main/scala/com/github/jonnylaw/model/X.scala of the reference.
"""
import bisect
import json
import math
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------
# parameters: the transforms of the smart constructors and of the SDE constructors
# ---------------------------------------------------------------------------------------------
def logistic(x):
    return 1.0 / (1.0 + math.exp(-x))


def repeat(dim, v):  # buildParamRepeat, model/Sde.scala:177-179
    return [v[i % len(v)] for i in range(dim)]


def ou_leaf(dim, m0, c0, phi, mu, sigma):
    """SdeParameter.ouParameter (model/SdeParameters.scala:202-205: c0, sigma -> log; phi -> logistic)
    followed by OuProcess (model/Sde.scala:133-137: c0, sigma -> exp; phi -> logistic again)."""
    return dict(kind="ou", dim=dim, raw=dict(m0=m0, c0=c0, phi=phi, mu=mu, sigma=sigma),
                m0=repeat(dim, m0), c0=[math.exp(math.log(v)) for v in repeat(dim, c0)],
                phi=[logistic(logistic(v)) for v in repeat(dim, phi)], mu=repeat(dim, mu),
                sigma=[math.exp(math.log(v)) for v in repeat(dim, sigma)])


def bm_leaf(dim, m0, c0, sigma):
    """SdeParameter.brownianParameter (:190-193) + BrownianMotion (model/Sde.scala:99-102)."""
    return dict(kind="bm", dim=dim, raw=dict(m0=m0, c0=c0, sigma=sigma), m0=repeat(dim, m0),
                c0=[math.exp(math.log(v)) for v in repeat(dim, c0)],
                sigma=[math.exp(math.log(v)) for v in repeat(dim, sigma)])


def genbm_leaf(dim, m0, c0, mu, sigma):
    """SdeParameter.genBrownianParameter (:176-181) + GenBrownianMotion (model/Sde.scala:70-73)."""
    return dict(kind="genbm", dim=dim, raw=dict(m0=m0, c0=c0, mu=mu, sigma=sigma), m0=repeat(dim, m0),
                c0=[math.exp(math.log(v)) for v in repeat(dim, c0)], mu=repeat(dim, mu),
                sigma=[math.exp(math.log(v)) for v in repeat(dim, sigma)])


# ---------------------------------------------------------------------------------------------
# a3-a7: initial state, exact transitions, f, dataLikelihood
# ---------------------------------------------------------------------------------------------
def initial_state(leaf, z):  # model/Sde.scala:75-80,104-108,152-156: m0 + sqrt(c0) * z
    return [leaf["m0"][k] + math.sqrt(leaf["c0"][k]) * z[k] for k in range(leaf["dim"])]


def step_exact(leaf, x, dt, z):
    out = []
    for k in range(leaf["dim"]):
        if leaf["kind"] == "bm":  # model/Sde.scala:114-123: x + sqrt(sigma*dt) * z
            out.append(x[k] + math.sqrt(leaf["sigma"][k] * dt) * z[k])
        elif leaf["kind"] == "genbm":  # :86-95: x + mu*dt + sqrt(sigma*dt) * z
            out.append(x[k] + leaf["mu"][k] * dt + math.sqrt(leaf["sigma"][k] * dt) * z[k])
        else:  # OU :139-150
            phi, mu, sigma = leaf["phi"][k], leaf["mu"][k], leaf["sigma"][k]
            variance = (sigma * sigma / (2.0 * phi)) * (1.0 - math.exp(-2.0 * phi * dt))
            mean = mu + (x[k] - mu) * math.exp(-phi * dt)
            out.append(math.sqrt(variance) * z[k] + mean)
    return out


def step_euler(leaf, x, dt, z):  # model/Sde.scala:30-43 with drift/diffusion of :82-84,:110-112,:158-162
    out = []
    for k in range(leaf["dim"]):
        if leaf["kind"] == "bm":
            drift = 1.0  # sic, model/Sde.scala:110
        elif leaf["kind"] == "genbm":
            drift = leaf["mu"][k]
        else:
            drift = leaf["phi"][k] * (leaf["mu"][k] - x[k])
        out.append(x[k] + drift * dt + leaf["sigma"][k] * (math.sqrt(dt) * z[k]))
    return out


def f_leaf(mleaf, x, t):
    if mleaf["f"] == "first":  # model/Model.scala:184,250,271,328,366
        return x[0]
    # SeasonalModel.buildF, model/Model.scala:217-225
    omega = 2.0 * math.pi / mleaf["period"]
    acc = 0.0
    for a in range(1, mleaf["harmonics"] + 1):
        acc += math.cos(omega * a * t) * x[2 * (a - 1)]
        acc += math.sin(omega * a * t) * x[2 * (a - 1) + 1]
    return acc


def log_density(model, gamma, y):
    kind = model["obs"]
    if kind == "poisson":  # Poisson(exp gamma).logProbabilityOf(y.toInt), model/Model.scala:269-273
        lam, k = math.exp(gamma), int(y)
        return -lam + k * math.log(lam) - math.lgamma(k + 1.0)
    if kind == "negbin":  # model/Model.scala:186-195
        size, mu, k = math.exp(model["scale"]), math.exp(gamma), int(y)
        return (math.lgamma(size + k) - math.lgamma(k + 1.0) - math.lgamma(size) + size * math.log(size / (mu + size)) +
                k * math.log(mu / (mu + size)))
    if kind == "normal":  # Gaussian(gamma, exp(scale)).logPdf(y), :227-233,:252-258
        sd = math.exp(model["scale"])
        dd = (y - gamma) / sd
        return -dd * dd / 2.0 - math.log(math.sqrt(2.0 * math.pi)) - math.log(sd)
    if kind == "bernoulli":  # model/Model.scala:318-336
        p = 1.0 if gamma > 6 else (0.0 if gamma < -6 else 1.0 / (1.0 + math.exp(-gamma)))
        if y == 1.0:
            return -1e99 if p == 0.0 else math.log(p)
        return -1e99 if p == 1.0 else math.log(1.0 - p)
    if kind == "student_t":  # 1/v * StudentsT(df).logPdf((y - eta)/v), model/Model.scala:154-160 (the 1/v factor is the reference's)
        v, df = math.exp(model["scale"]), float(model["df"])
        x = (y - gamma) / v
        lp = (math.lgamma((df + 1.0) / 2.0) - math.lgamma(df / 2.0) - 0.5 * math.log(math.pi * df) -
              (df + 1.0) / 2.0 * math.log(1.0 + x * x / df))
        return 1.0 / v * lp
    if kind == "zip":  # model/Model.scala:298-306
        v, k = model["scale"], int(y)
        p = math.exp(v) / (1.0 + math.exp(v))
        if k == 0:
            return math.log(p + (1.0 - p) * math.exp(-math.exp(gamma)))
        return -math.log(1.0 + math.exp(v)) + k * gamma - math.exp(gamma) - math.lgamma(k + 1.0)
    if kind == "beta":  # new Beta(exp(-gamma), 1.0).logPdf(y), model/Model.scala:349-352
        a, b = math.exp(-gamma), 1.0
        return (a - 1.0) * math.log(y) + (b - 1.0) * math.log(1.0 - y) - (math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b))
    raise ValueError(kind)


# ---------------------------------------------------------------------------------------------
# a8-a9: weights, ll, ESS, the TreeMap ECDF and the three resamplers
# ---------------------------------------------------------------------------------------------
def fold_sum(v):  # foldLeft(0.0)(_ + _)
    acc = 0.0
    for a in v:
        acc = acc + a
    return acc


def tree_ecdf(w):
    """Resampling.treeEcdf (model/Resampling.scala:52-58): normalise (sequential total), scanLeft,
    TreeMap ++ (cumulative -> item): a repeated key is overwritten, the LAST item survives."""
    total = fold_sum(w)
    wn = [a / total for a in w]
    cum, acc = [], 0.0
    for a in wn:
        acc = acc + a
        cum.append(acc)
    m = {}
    for j, c in enumerate(cum):
        m[c] = j
    keys = sorted(m)
    return keys, m


def find_all(ks, keys, m):
    """findAllInTreeMap (:36-46): for each k the first entry with key >= k (m.from(k).head).
    Where the reference would throw (no such key: k above the last key by rounding) the last
    particle is returned and the case is reported."""
    out, clamped = [], 0
    for k in ks:
        i = bisect.bisect_left(keys, k)
        if i == len(keys):
            clamped += 1
            i = len(keys) - 1
        out.append(m[keys[i]])
    return out, clamped


def first_index(w, ks):
    """NOT the reference: the textbook inverse CDF on the same cumulative sums, first index whose cumulative
    weight reaches k (no key is overwritten).  Golden values for the library's CSSM_TIE_FIRST option."""
    total = fold_sum(w)
    cum, acc = [], 0.0
    for a in w:
        acc = acc + a / total
        cum.append(acc)
    return [min(bisect.bisect_left(cum, k), len(w) - 1) for k in ks]


def systematic(w, u):  # :63-72
    n = len(w)
    keys, m = tree_ecdf(w)
    return find_all([(u + i) / n for i in range(n)], keys, m)


def stratified(w, us):  # :78-86
    n = len(w)
    keys, m = tree_ecdf(w)
    return find_all([(i + us[i]) / n for i in range(n)], keys, m)


def multinomial(w, us):
    """:92-96 with Breeze Multinomial.draw (first-draw path): walk subtracting the weights from
    u * sum until <= 0."""
    total = fold_sum(w)
    out = []
    for u in us:
        prob = u * total
        i = 0
        while True:
            prob = prob - w[i]
            if prob <= 0 or i == len(w) - 1:
                break
            i += 1
        out.append(i)
    return out


def ll_ess(logw):
    """stepFilter, model/ParticleFilter.scala:124-128; effectiveSampleSize :431-434; mean :522-524."""
    mx = max(logw)
    w1 = [math.exp(a - mx) for a in logw]
    incr = mx + math.log(fold_sum(w1) / len(w1))
    total = fold_sum(w1)
    s2 = fold_sum([(a / total) * (a / total) for a in w1])
    inv = math.floor(1.0 / s2)
    return mx, w1, incr, int(inv)


# ---------------------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------------------
def model_c1():
    return dict(name="c1", obs="poisson", scale=None,
                leaves=[dict(f="first", sde=ou_leaf(1, [1.0], [0.5], [0.2], [1.5], [0.05]))])


def model_c2():
    return dict(name="c2", obs="poisson", scale=None,
                leaves=[dict(f="first", sde=ou_leaf(1, [1.0], [0.5], [0.2], [1.5], [0.05])),
                        dict(f="seasonal", period=24, harmonics=3, sde=ou_leaf(6, [0.1], [1.0], [0.4], [0.1], [0.5]))])


def model_c4():
    return dict(name="c4", obs="negbin", scale=2.0,
                leaves=[dict(f="first", sde=bm_leaf(1, [0.0], [1.0], [0.01])),
                        dict(f="first", sde=genbm_leaf(1, [0.0], [1.0], [0.01], [0.01]))])


def model_c5():
    return dict(name="c5", obs="normal", scale=0.0,
                leaves=[dict(f="first", sde=ou_leaf(1, [1.0], [0.5], [0.2], [1.5], [0.05])),
                        dict(f="seasonal", period=24, harmonics=3, sde=ou_leaf(6, [0.1], [1.0], [0.4], [0.1], [0.5]))])


def model_bernoulli():
    return dict(name="bernoulli", obs="bernoulli", scale=None,
                leaves=[dict(f="first", sde=bm_leaf(2, [0.0, 0.5], [1.0], [0.3]))])


def model_student():
    return dict(name="student_t", obs="student_t", scale=-0.7, df=5,
                leaves=[dict(f="first", sde=ou_leaf(1, [0.5], [0.4], [0.3], [0.2], [0.3]))])


def model_zip():
    return dict(name="zip", obs="zip", scale=-1.2,
                leaves=[dict(f="first", sde=ou_leaf(1, [1.0], [0.5], [0.2], [1.5], [0.05])),
                        dict(f="seasonal", period=12, harmonics=1, sde=ou_leaf(2, [0.1], [0.5], [0.4], [0.0], [0.3]))])


def model_beta():
    return dict(name="beta", obs="beta", scale=2.0,
                leaves=[dict(f="first", sde=bm_leaf(1, [0.3], [0.2], [0.05]))])


def dim(model):
    return sum(l["sde"]["dim"] for l in model["leaves"])


def split(model, flat):
    out, k = [], 0
    for l in model["leaves"]:
        out.append(flat[k:k + l["sde"]["dim"]])
        k += l["sde"]["dim"]
    return out


def model_f(model, flat, t):  # composed f = f1 + f2, model/Model.scala:122-128
    parts = split(model, flat)
    g = None
    for l, x in zip(model["leaves"], parts):
        v = f_leaf(l, x, t)
        g = v if g is None else g + v
    return g


def filter_case(model, N, T, seed, missing=(), euler=False, big_y=None):
    rng = random.Random(seed)
    d = dim(model)
    gauss = lambda: rng.gauss(0.0, 1.0)
    z0 = [[gauss() for _ in range(d)] for _ in range(N)]
    xs = []
    for i in range(N):
        parts, k = [], 0
        for l in model["leaves"]:
            dd = l["sde"]["dim"]
            parts += initial_state(l["sde"], z0[i][k:k + dd])
            k += dd
        xs.append(parts)
    case = dict(model=model, N=N, d=d, euler=euler, z0=z0, x0=[list(x) for x in xs], t0=0.0, steps=[])
    t_prev, ll = 0.0, 0.0
    ess = N
    for s in range(T):
        t = 0.1 * s if s > 0 else 0.0  # first datum at t0: dt = 0 (model/ParticleFilter.scala:138)
        dt = t - t_prev
        z = [[gauss() for _ in range(d)] for _ in range(N)]
        xp = []
        for i in range(N):
            parts, k = [], 0
            for l, x in zip(model["leaves"], split(model, xs[i])):
                dd = l["sde"]["dim"]
                stepf = step_euler if euler else step_exact
                parts += stepf(l["sde"], x, dt, z[i][k:k + dd])
                k += dd
            xp.append(parts)
        has_obs = s not in missing
        # a synthetic observation near the cloud (or an extreme one, to make weights degenerate)
        g_mean = fold_sum([model_f(model, x, t) for x in xp]) / N
        if model["obs"] in ("poisson", "negbin"):
            y = float(max(0, int(round(math.exp(g_mean) + rng.choice([-1, 0, 1, 2])))))
        elif model["obs"] == "zip":
            y = 0.0 if s % 2 == 0 else float(max(0, int(round(math.exp(g_mean)))))
        elif model["obs"] in ("normal", "student_t"):
            y = g_mean + 0.3 * gauss()
        elif model["obs"] == "beta":
            y = min(max(0.5 + 0.2 * gauss(), 0.05), 0.95)
        else:
            y = 1.0 if rng.random() < 0.5 else 0.0
        if big_y is not None and s in big_y:
            y = big_y[s]
        step = dict(t=t, has_obs=has_obs, y=y, z=z, x_prop=[list(x) for x in xp])
        if not has_obs:  # propagated cloud, ll and ess unchanged (:121)
            xs = xp
            step.update(ll=ll, ess=ess)
        else:
            logw = [log_density(model, model_f(model, x, t), y) for x in xp]
            mx, w1, incr, e = ll_ess(logw)
            ll, ess = ll + incr, e
            u_sys = rng.random()
            u_n = [rng.random() for _ in range(N)]
            a_sys, c_sys = systematic(w1, u_sys)
            a_str, c_str = stratified(w1, u_n)
            a_mul = multinomial(w1, u_n)
            step.update(logw=logw, max=mx, w1=w1, ll_incr=incr, ll=ll, ess=ess, u_sys=u_sys, u_n=u_n, anc_systematic=a_sys,
                        anc_stratified=a_str, anc_multinomial=a_mul, clamped=[c_sys, c_str])
            xs = [xp[j] for j in a_sys]  # the cases continue with systematic resampling
        case["steps"].append(step)
        t_prev = t
    return case


def resample_cases():
    rng = random.Random(99)
    out = []
    weights = {
        "unit": [1.0] * 9,
        "ties_zero_runs": [0.5, 0.0, 0.0, 0.25, 0.0, 0.25, 0.0],
        "leading_zeros": [0.0, 0.0, 1.0, 3.0],
        "one_heavy": [1e-30] * 5 + [1.0] + [1e-30] * 6,
        "vanishing_after_big": [1.0, 1e-17, 1e-17, 1e-17, 1.0, 1e-18],
        "random": [math.exp(2.0 * rng.gauss(0, 1)) for _ in range(37)],
        "single": [0.7],
    }
    for name, w in weights.items():
        n = len(w)
        for rep in range(3):
            u = rng.random()
            us = [rng.random() for _ in range(n)]
            if rep == 2:
                u, us = 0.0, [0.0] * n
            a_sys, c1 = systematic(w, u)
            a_str, c2 = stratified(w, us)
            out.append(dict(name=name, w=w, u=u, us=us, anc_systematic=a_sys, anc_stratified=a_str,
                            anc_multinomial=multinomial(w, us), clamped=[c1, c2],
                            anc_systematic_first=first_index(w, [(u + i) / n for i in range(n)]),
                            anc_stratified_first=first_index(w, [(i + us[i]) / n for i in range(n)])))
    return out


def lgcp_case():
    """FilterLgcp.calcWeight / stepFilter (model/ParticleFilter.scala:184-226): n = ceil(dt / 10^-p)
    exact sub-steps, hazard over the n post-step states at times t + i*delta, log-weight
    f(x_n, t) - hazard; dt == 0 gives f - f."""
    rng = random.Random(5)
    leaf = bm_leaf(1, [0.0], [1.0], [0.01])
    precision, N = 2, 6
    delta = 10.0 ** (-precision)
    x = [[0.3 * rng.gauss(0, 1)] for _ in range(N)]
    x_start = [list(v) for v in x]
    steps = []
    t_prev = 0.0
    for t in (0.0, 0.035, 0.06):
        dt = t - t_prev
        n_sub = 0 if dt == 0 else int(math.ceil(dt / delta))
        z = [[[rng.gauss(0, 1)] for _ in range(N)] for _ in range(n_sub)]
        xp, logw = [], []
        for i in range(N):
            xi = list(x[i])
            hz, time = 0.0, t
            for s in range(n_sub):
                xi = step_exact(leaf, xi, delta, z[s][i])
                time = time + delta
                hz = hz + math.exp(xi[0]) * delta
            g = xi[0]
            logw.append(g - g if n_sub == 0 else g - hz)
            xp.append(xi)
        mx, w1, incr, ess = ll_ess(logw)
        us = [rng.random() for _ in range(N)]
        a_str, _ = stratified(w1, us)
        steps.append(dict(t=t, n_sub=n_sub, z=z, x_prop=xp, logw=logw, ll_incr=incr, ess=ess, u_n=us, anc_stratified=a_str))
        x = [xp[j] for j in a_str]
        t_prev = t
    return dict(leaf=leaf, precision=precision, N=N, x_start=x_start, steps=steps)


def main():
    cases = [
        filter_case(model_c1(), 16, 4, 1),
        filter_case(model_c2(), 12, 3, 2),
        filter_case(model_c4(), 10, 4, 3, missing=(2,)),
        filter_case(model_c5(), 9, 3, 4, big_y={2: 40.0}),
        filter_case(model_bernoulli(), 11, 3, 5),
        filter_case(model_c4(), 7, 3, 6, euler=True),
        filter_case(model_student(), 9, 3, 7),
        filter_case(model_zip(), 10, 4, 8),
        filter_case(model_beta(), 8, 3, 9),
    ]
    json.dump(cases, open(os.path.join(HERE, "filter_steps.json"), "w"), indent=None, separators=(",", ":"))
    json.dump(resample_cases(), open(os.path.join(HERE, "resampling.json"), "w"), indent=None, separators=(",", ":"))
    json.dump(lgcp_case(), open(os.path.join(HERE, "lgcp_steps.json"), "w"), indent=None, separators=(",", ":"))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".json")))


if __name__ == "__main__":
    main()
