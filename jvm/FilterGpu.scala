// FilterGpu.scala -- the Scala side of the drop-in.  NOT COMPILED IN THIS IMAGE (no JVM / sbt / Breeze jars).
// It lives in package com.github.jonnylaw.model on purpose: the concrete model classes are
// `private final case class` (model/Model.scala:168,204,241,266,315,363), so a descriptor can only be
// built from inside the package (SURVEY.md 8b).
package com.github.jonnylaw.model

import akka.NotUsed
import akka.stream.scaladsl.Flow
import breeze.linalg.DenseVector
import breeze.stats.distributions.Rand
import cats.data.Reader
import com.github.jonnylaw.gpu.CssmNative

/** Flat description of a parameterised model: what include/cssm.h calls cssm_model_desc_t. */
final case class GpuDesc(kinds: Array[Int], params: Array[Double], obsKind: Int, scale: Option[Double],
  stepMode: Int = 0, precision: Int = 0, obsDf: Int = 0)

object GpuDesc {
  private def leafOf(sde: Sde): (Array[Int], Array[Array[Double]]) = sde match {
    case BrownianMotion(_, d) => val p = sde.asInstanceOf[BrownianMotion].params
      (Array(0, d), Array(p.m0.toArray, p.c0.toArray, Array.fill(d)(0.0), Array.fill(d)(0.0), p.sigma.toArray))
    case GenBrownianMotion(_, d) => val p = sde.asInstanceOf[GenBrownianMotion].params
      (Array(1, d), Array(p.m0.toArray, p.c0.toArray, Array.fill(d)(0.0), p.mu.toArray, p.sigma.toArray))
    case OuProcess(_, d) => val p = sde.asInstanceOf[OuProcess].params
      (Array(2, d), Array(p.m0.toArray, p.c0.toArray, p.phi.toArray, p.mu.toArray, p.sigma.toArray))
  }
  /** (obsKind, fKind, period, harmonics, scale, sde) of one un-composed model */
  private def single(m: Model): (Int, Int, Int, Int, Option[Double], Sde) = m match {
    case PoissonModel(s, p)            => (0, 0, 0, 0, p.scale, s)
    case NegativeBinomialModel(s, p)   => (1, 0, 0, 0, p.scale, s)
    case LinearModel(s, p)             => (2, 0, 0, 0, p.scale, s)
    case SeasonalModel(per, h, s, p)   => (2, 1, per, h, p.scale, s)
    case BernoulliModel(s, p)          => (3, 0, 0, 0, p.scale, s)
    case LogGaussianCox(s, p)          => (4, 0, 0, 0, p.scale, s)
    case StudentsTModel(s, _, p)       => (5, 0, 0, 0, p.scale, s)   // df travels in GpuDesc.obsDf
    case ZeroInflatedPoisson(s, p)     => (6, 0, 0, 0, p.scale, s)
    case BetaModel(s, p)               => (7, 0, 0, 0, p.scale, s)
  }
  /** leaves in Tree.flatten order; `models` is the list of un-composed models, left to right */
  def apply(models: List[Model], precision: Int = 0): GpuDesc = {
    val singles = models.map(single)
    val leaves = singles.map { case (_, f, per, h, _, s) => val (k, p) = leafOf(s); (Array(k(0), k(1), f, per, h), p) }
    val d = leaves.map(_._1(1)).sum
    val params = (0 until 5).toArray.flatMap(i => leaves.flatMap(_._2(i)))   // m0 | c0 | phi | mu | sigma
    val df = models.head match { case StudentsTModel(_, df, _) => df; case _ => 0 }
    GpuDesc(leaves.flatMap(_._1).toArray, params, singles.head._1, singles.head._5, 0, precision, df)
  }
}

/** trait ParticleFilter[State] (model/ParticleFilter.scala:96-167) with the particle work on the GPU.
  * `PfState.particles` is a strict `Vector[State]` in the reference, and `Vector` cannot be subclassed, so a lazy view
  * is not possible: with `materialise = true` every returned PfState carries the cloud (an N x d copy from the device per
  * step, what the reference's own callers see); with `materialise = false` (default) `particles` is empty and the cloud
  * stays on the device -- read it with `particles()`, summarise it with `intervals` / `getMeanForecast` below.  ll and ess
  * are always filled. */
final case class FilterGpu(models: List[Model], mod: Model, resampleKind: Int, precision: Int = 0,
  dtype: Int = 0, device: Int = 0, seed: Long = 0L, materialise: Boolean = false) extends ParticleFilter[State] with AutoCloseable {

  private val desc = GpuDesc(models, precision)
  private var handle: Long = 0L
  private var n: Int = 0
  private val dims: List[Int] = models.map(_.sde.dimension)

  private def ensure(particles: Int): Long = {
    if (handle == 0L || n != particles) {
      close()
      handle = CssmNative.filterCreate(desc.kinds, desc.params, desc.obsKind, desc.scale.isDefined,
        desc.scale.getOrElse(0.0), desc.stepMode, desc.precision, desc.obsDf, particles.toLong, resampleKind, dtype, device,
        seed, 0L)
      n = particles
    }
    handle
  }
  def close(): Unit = if (handle != 0L) { CssmNative.filterDestroy(handle); handle = 0L }

  // the abstract members are only used by the CPU code paths we override
  def dataLikelihood(g: Gamma, y: Observation) = mod.dataLikelihood(g, y)
  def stepFunction(dt: TimeIncrement)(s: State): Rand[State] = mod.sde.stepFunction(dt)(s)
  def initialState = mod.sde.initialState
  def f(s: State, t: Time) = mod.f(s, t)
  def resample: Resample[State] = (p, w) => GpuResample(resampleKind, device)(p, w)

  /** the current (resampled) cloud, copied out of device memory */
  def particles(): Vector[State] = cloud()
  private def held(): Vector[State] = if (materialise) cloud() else Vector.empty
  private def cloud(): Vector[State] = {
    val d = dims.sum
    val flat = new Array[Double](d * n)
    CssmNative.filterGetParticles(handle, flat)              // [d][N]
    Vector.tabulate(n) { i =>
      var off = 0
      dims.map { dim => val v = DenseVector.tabulate(dim)(k => flat((off + k) * n + i)); off += dim; Tree.leaf(v): State }
        .reduceLeft(_ +++ _)
    }
  }

  override def initialiseState(particles: Int, t0: Time): PfState[State] = {
    CssmNative.filterInit(ensure(particles), t0)
    PfState(t0, None, held(), 0.0, particles)
  }
  override def stepFilter(s: PfState[State], y: Data): PfState[State] = {
    val ess = new Array[Int](1)
    val ll = CssmNative.filterStep(handle, y.t, y.observation.isDefined, y.observation.getOrElse(0.0), ess)
    PfState(y.t, y.observation, held(), ll, ess(0))
  }
  override def llFilter(data: Vector[Data], particles: Int): LogLikelihood =
    CssmNative.filterLl(ensure(particles), data.map(_.t).toArray, data.map(_.observation.getOrElse(0.0)).toArray,
      data.map(d => (if (d.observation.isDefined) 1 else 0).toByte).toArray)
  override def filter(data: Vector[Data], particles: Int): (LogLikelihood, Vector[StateSpace[State]]) = {
    val d = dims.sum
    val states = new Array[Double]((data.size + 1) * d)
    val ll = CssmNative.filterRun(ensure(particles), data.map(_.t).toArray, data.map(_.observation.getOrElse(0.0)).toArray,
      data.map(x => (if (x.observation.isDefined) 1 else 0).toByte).toArray, states)
    val times = data.minBy(_.t).t +: data.map(_.t)
    (ll, times.zipWithIndex.map { case (t, s) =>
      var off = s * d
      StateSpace(t, dims.map { dim => val v = DenseVector(states.slice(off, off + dim)); off += dim; Tree.leaf(v): State }
        .reduceLeft(_ +++ _)) })
  }
  private def toState(v: Array[Double], from: Int): State = {
    var off = from
    dims.map { dim => val x = DenseVector(v.slice(off, off + dim)); off += dim; Tree.leaf(x): State }.reduceLeft(_ +++ _)
  }
  /** ParticleFilter.getIntervals (model/ParticleFilter.scala:415-424) of the handle's current cloud, on the device */
  def intervals(s: PfState[State], interval: Double = 0.975): PfOut[State] = {
    val d = dims.sum
    val o = new Array[Double](3 * d + 2)
    CssmNative.filterIntervals(handle, s.t, interval, d, o)
    val (lo, up) = (mod.link(o(3 * d)), mod.link(o(3 * d + 1)))
    PfOut(s.t, s.observation, mod.link(mod.f(toState(o, 0), s.t)), CredibleInterval(math.min(lo, up), math.max(lo, up)),
      toState(o, 0), (0 until d).map(k => CredibleInterval(o(d + k), o(2 * d + k))))
  }
  /** ParticleFilter.getMeanForecast (model/ParticleFilter.scala:394-412) of the handle's current cloud, on the device:
    * only the 3(d + 2) summary numbers cross the boundary.  `chain` continues from the previous forecast cloud
    * (SimulateData.forecast, model/Data.scala:202-217). */
  def getMeanForecast(t: Time, interval: Double, chain: Boolean = false): ForecastOut[State] = {
    val d = dims.sum
    val o = new Array[Double](3 * d + 6)
    CssmNative.filterForecast(handle, t, interval, chain, d, o)
    ForecastOut(t, o(3 * d + 3), CredibleInterval(o(3 * d + 4), o(3 * d + 5)), o(3 * d), CredibleInterval(o(3 * d + 1), o(3 * d + 2)),
      toState(o, 0), (0 until d).map(k => CredibleInterval(o(d + k), o(2 * d + k))))
  }
  /** FilterInterpolate (model/ParticleFilter.scala:273-311): enable before initialiseState; `paths(idx)` returns the paths
    * of the given particles, newest state first like the reference's List[State]. */
  def enablePaths(maxSteps: Int): Unit = CssmNative.filterPathsEnable(handle, maxSteps.toLong)
  def paths(idx: Array[Int]): Vector[List[State]] = {
    val d = dims.sum
    val len = CssmNative.filterPathsLen(handle).toInt + 1
    val flat = new Array[Double](idx.length * len * d)
    CssmNative.filterGetPaths(handle, idx, flat)
    Vector.tabulate(idx.length)(p => List.tabulate(len)(s => toState(flat, (p * len + (len - 1 - s)) * d)))
  }
  override def filterStream(t0: Time, particles: Int): Flow[Data, PfState[State], NotUsed] =
    Flow[Data].scan(initialiseState(particles, t0))(stepFilter)           // same shape as model/ParticleFilter.scala:163-166
}

/** Resample[A] backed by cssm_resample: uniforms from scala.util.Random as in model/Resampling.scala:66,83 */
final case class GpuResample(kind: Int, device: Int = 0) {
  def apply[A](particles: Vector[A], weights: Vector[LogLikelihood]): Vector[A] = {
    val n = weights.size
    val u = if (kind == 0) Array(scala.util.Random.nextDouble) else Array.fill(n)(scala.util.Random.nextDouble)
    val anc = new Array[Int](n)
    CssmNative.resample(kind, weights.toArray, u, anc, device)
    anc.toVector.map(particles(_))
  }
}

object FilterGpu {
  /** BootstrapFilter for PMMH (model/PMMH.scala:58, examples/DetermineParameters.scala:67-72) over ONE device handle
    * for the life of the chain: the observations are uploaded once (filterLoadSeries), every proposal re-parameterises
    * the handle (filterSetParams: same shapes, no reallocation, the per-observation constants are rebuilt and sent
    * asynchronously from pinned memory) and runs the resident series (filterLlResident); `state._2.last`, all that
    * mhStep reads of the filtered states (model/PMMH.scala:76), is one particle of the final cloud (filterSampleOne).
    * `build` turns Parameters into (list of un-composed models, composed model).  Close it when the chain is done. */
  final class Bootstrap(build: Parameters => (List[Model], Model), data: Vector[Data], resampleKind: Int, n: Int,
      precision: Int = 0, dtype: Int = 0, device: Int = 0, seed: Long = 0L, streamId: Long = 0L) extends AutoCloseable {
    private var handle: Long = 0L
    private val t = data.map(_.t).toArray
    private val y = data.map(_.observation.getOrElse(0.0)).toArray
    private val ho = data.map(d => (if (d.observation.isDefined) 1 else 0).toByte).toArray
    private val tLast = data.last.t

    private def eval(p: Parameters): (LogLikelihood, Vector[StateSpace[State]]) = {
      val (ms, _) = build(p)
      val desc = GpuDesc(ms, precision)
      if (handle == 0L) {
        handle = CssmNative.filterCreate(desc.kinds, desc.params, desc.obsKind, desc.scale.isDefined, desc.scale.getOrElse(0.0),
          desc.stepMode, desc.precision, desc.obsDf, n.toLong, resampleKind, dtype, device, seed, streamId)
        CssmNative.filterLoadSeries(handle, t, y, ho)
      } else {
        CssmNative.filterSetParams(handle, desc.kinds, desc.params, desc.obsKind, desc.scale.isDefined,
          desc.scale.getOrElse(0.0), desc.stepMode, desc.precision, desc.obsDf)
      }
      val ll = CssmNative.filterLlResident(handle)
      val dims = ms.map(_.sde.dimension)
      val x = new Array[Double](dims.sum)
      CssmNative.filterSampleOne(handle, x)
      var off = 0
      val state = dims.map { dim => val v = DenseVector(x.slice(off, off + dim)); off += dim; Tree.leaf(v): State }.reduceLeft(_ +++ _)
      (ll, Vector(StateSpace(tLast, state)))
    }
    /** what MetropolisHastings.pf expects */
    val reader: BootstrapFilter[Parameters, StateSpace[State]] = Reader(eval)
    def close(): Unit = if (handle != 0L) { CssmNative.filterDestroy(handle); handle = 0L }
  }
  def bootstrap(build: Parameters => (List[Model], Model), data: Vector[Data], resampleKind: Int, n: Int,
      precision: Int = 0, dtype: Int = 0, device: Int = 0, seed: Long = 0L, streamId: Long = 0L): Bootstrap =
    new Bootstrap(build, data, resampleKind, n, precision, dtype, device, seed, streamId)
}
