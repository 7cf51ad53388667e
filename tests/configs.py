"""The BASELINE.json configurations as host-side models, with the
reference's example parameter values (examples/Simulation.scala:16,24,64-67)."""
import numpy as np

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import Model, Sde, SdeParameter, Parameters, _abi

SYS, STRAT, MULTI = _abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED, _abi.RESAMPLE_MULTINOMIAL


def ou1():
    return SdeParameter.ouParameter([1.0], [0.5], [0.2], [1.5], [0.05])


def ou6():
    return SdeParameter.ouParameter([0.1], [1.0], [0.4], [0.1], [0.5])


def c1():
    """Poisson observations, OU latent state (d = 1)."""
    return Model.poisson(Sde.ouProcess(1))(Parameters(None, ou1()))


def c2():
    """Poisson + seasonal(24, 3) + OU (d = 7)."""
    m = Model.poisson(Sde.ouProcess(1)) | Model.seasonal(24, 3, Sde.ouProcess(6))
    return m(Parameters(None, ou1()) | Parameters(None, ou6()))


def c3(precision=2):
    """log-Gaussian Cox process, Brownian-motion latent state (d = 1)."""
    m = Model.lgcp(Sde.brownianMotion(1))(Parameters(None, SdeParameter.brownianParameter([0.0], [1.0], [0.01])))
    return cs.model.Model(m.leaves, m.step_mode, precision)


def c4_unparam():
    return Model.negativeBinomial(Sde.brownianMotion(1)) | Model.linear(Sde.genBrownianMotion(1))


def c4_params():
    return (Parameters(2.0, SdeParameter.brownianParameter([0.0], [1.0], [0.01])) |
            Parameters(None, SdeParameter.genBrownianParameter([0.0], [1.0], [0.01], [0.01])))


def c4():
    """negative binomial + linear trend (d = 2)."""
    return c4_unparam()(c4_params())


def c5():
    """Normal + seasonal(24, 3) + OU (d = 7)."""
    m = Model.linear(Sde.ouProcess(1)) | Model.seasonal(24, 3, Sde.ouProcess(6))
    return m(Parameters(0.0, ou1()) | Parameters(None, ou6()))


def bernoulli_bm():
    return Model.bernoulli(Sde.brownianMotion(2))(Parameters(None, SdeParameter.brownianParameter([0.0, 0.5], [1.0], [0.3])))


def normal_genbm():
    return Model.linear(Sde.genBrownianMotion(1))(Parameters(-0.5, SdeParameter.genBrownianParameter([0.2], [0.8], [0.05], [0.2])))


def student_ou():
    """Student-t observations (df = 5), OU latent state (model/Model.scala:144-162)."""
    return Model.studentsT(Sde.ouProcess(1), 5)(Parameters(-0.7, SdeParameter.ouParameter([0.5], [0.4], [0.3], [0.2], [0.3])))


def zip_seasonal():
    """zero-inflated Poisson + seasonal (model/Model.scala:281-309)."""
    m = Model.zeroInflatedPoisson(Sde.ouProcess(1)) | Model.seasonal(12, 1, Sde.ouProcess(2))
    return m(Parameters(-1.2, ou1()) | Parameters(None, SdeParameter.ouParameter([0.1], [0.5], [0.4], [0.0], [0.3])))


def beta_bm():
    """Beta observations, Brownian-motion latent state (model/Model.scala:339-353)."""
    return Model.beta(Sde.brownianMotion(1))(Parameters(2.0, SdeParameter.brownianParameter([0.3], [0.2], [0.05])))


def poisson_big():
    """Poisson observations with a mean around 60 (the transformed-rejection branch of the forecast sampler)."""
    return Model.poisson(Sde.ouProcess(1))(Parameters(None, SdeParameter.ouParameter([4.0], [0.05], [0.2], [4.1], [0.1])))


ALL = {"c1": c1, "c2": c2, "c4": c4, "c5": c5, "bernoulli": bernoulli_bm, "normal_genbm": normal_genbm,
       "student_t": student_ou, "zip": zip_seasonal, "beta": beta_bm}
EXTRA = {"poisson_big": poisson_big}
