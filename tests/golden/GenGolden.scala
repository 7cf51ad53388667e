// GenGolden.scala -- regenerate golden vectors for tests/golden/ from the UNMODIFIED reference.
//
// WHY.  The oracle (oracle/cssm_oracle.cpp) is pinned to vectors written by tests/golden/make_golden.py, a second
// restatement of the Scala by the same authors; the build image has no JVM, so nothing in this repository has ever been
// compared with output of jonnylaw/ComposableStateSpaceModels itself ("parity unpinned", DESIGN.md section 2).  This
// program closes that gap on any machine with sbt: it runs the reference's own `initialiseState`, `stepFilter` and
// resamplers under seeded generators, records the random numbers they consumed, and writes `filter_steps.ref.json` and
// `resampling.ref.json` in the schema of filter_steps.json / resampling.json.  With those two files in tests/golden/,
// `pytest tests/test_golden.py` checks the oracle against REAL reference output (test_*_against_reference_output; skipped
// while the files are absent) and the GPU suite inherits the pin through the oracle.
//
// STATUS.  Written against the reference sources (file:line cited below); NOT compiled here (no scalac in the image).
//
// HOW (from a checkout of the reference; Scala / sbt versions as its build.sbt says):
//   cp <this repo>/tests/golden/GenGolden.scala src/main/scala/com/github/jonnylaw/model/GenGolden.scala
//   sbt "runMain com.github.jonnylaw.model.GenGolden <this repo>/tests/golden"
// The file lives in package com.github.jonnylaw.model so that it sees the package's types without imports.
//
// HOW THE NOISE IS RECORDED WITHOUT TOUCHING THE REFERENCE.  Every normal the filter draws comes from Breeze's global
// basis -- `DenseVector.rand(dim, rand.gaussian(0, 1))` with the default `rand = Rand` (model/Sde.scala:78,91,106,119,146,
// 154) = `Rand.generator.nextGaussian`, in the order: particle by particle (Vector.fill / Vector.map), leaves left to right
// (model/Sde.scala:206-209,223-229), components in order.  The uniforms of systematic / stratified resampling come from
// the global `scala.util.Random` (model/Resampling.scala:66,83).  Both generators are seeded before each call, and a SECOND
// generator of the same class with the same seed is read in the same order: that yields the consumed numbers exactly.
// Each stage (propagate :118, weight :123, max / w1 :124-125, resample :126) is evaluated with the reference's OWN
// functions, and the composition is asserted equal to what `stepFilter` itself returns from the same seeds.
package com.github.jonnylaw.model

import java.io.PrintWriter

import breeze.stats.distributions.Rand
import cats.implicits._
import org.apache.commons.math3.random.MersenneTwister

object GenGolden {
  // ---- tiny JSON writer (17 significant digits: doubles round-trip) -------------------------------------------------
  private def num(x: Double): String = if (x.isNaN || x.isInfinite) "null" else "%.17g".format(x)
  private def arr(xs: Seq[Double]): String = xs.map(num).mkString("[", ", ", "]")
  private def iarr(xs: Seq[Int]): String = xs.mkString("[", ", ", "]")
  private def arr2(xs: Seq[Seq[Double]]): String = xs.map(arr).mkString("[", ", ", "]")
  private def flat(s: State): Seq[Double] = s.flatten.flatMap(_.data.toSeq)   // Tree.flatten: leaves left to right (model/Tree.scala:49-53)

  // ---- the models of BASELINE.json with the example values of examples/Simulation.scala:16,24,64-67 ------------------
  final case class Case(name: String, obs: String, scale: Option[Double], unparam: UnparamModel, params: Parameters,
    leavesJson: String, d: Int, ys: Vector[Double])

  private def ouJson(dim: Int, m0: Double, c0: Double, phi: Double, mu: Double, sigma: Double) =
    s"""{"kind": "ou", "dim": $dim, "raw": {"m0": [${num(m0)}], "c0": [${num(c0)}], "phi": [${num(phi)}], "mu": [${num(mu)}], "sigma": [${num(sigma)}]}}"""
  private def bmJson(m0: Double, c0: Double, sigma: Double) =
    s"""{"kind": "bm", "dim": 1, "raw": {"m0": [${num(m0)}], "c0": [${num(c0)}], "sigma": [${num(sigma)}]}}"""
  private def genbmJson(m0: Double, c0: Double, mu: Double, sigma: Double) =
    s"""{"kind": "genbm", "dim": 1, "raw": {"m0": [${num(m0)}], "c0": [${num(c0)}], "mu": [${num(mu)}], "sigma": [${num(sigma)}]}}"""

  private val ou1 = SdeParameter.ouParameter(1.0)(0.5)(0.2)(1.5)(0.05)        // model/SdeParameters.scala:202-205
  private val ou6 = SdeParameter.ouParameter(0.1)(1.0)(0.4)(0.1)(0.5)
  private val ou1J = ouJson(1, 1.0, 0.5, 0.2, 1.5, 0.05)
  private val ou6J = ouJson(6, 0.1, 1.0, 0.4, 0.1, 0.5)

  private val cases: List[Case] = List(
    Case("c1", "poisson", None, Model.poisson(Sde.ouProcess(1)), Parameters(None, ou1),
      s"""[{"f": "first", "sde": $ou1J}]""", 1, Vector(2.0, 1.0, 3.0, 0.0)),
    Case("c2", "poisson", None, Model.poisson(Sde.ouProcess(1)) |+| Model.seasonal(24, 3, Sde.ouProcess(6)),
      Parameters(None, ou1) |+| Parameters(None, ou6),
      s"""[{"f": "first", "sde": $ou1J}, {"f": "seasonal", "period": 24, "harmonics": 3, "sde": $ou6J}]""", 7, Vector(4.0, 2.0, 7.0, 1.0)),
    Case("c4", "negbin", Some(2.0), Model.negativeBinomial(Sde.brownianMotion(1)) |+| Model.linear(Sde.genBrownianMotion(1)),
      Parameters(Some(2.0), SdeParameter.brownianParameter(0.0)(1.0)(0.01)) |+|
        Parameters(None, SdeParameter.genBrownianParameter(0.0)(1.0)(0.01)(0.01)),
      s"""[{"f": "first", "sde": ${bmJson(0.0, 1.0, 0.01)}}, {"f": "first", "sde": ${genbmJson(0.0, 1.0, 0.01, 0.01)}}]""", 2,
      Vector(1.0, 0.0, 3.0, 2.0)),
    Case("c5", "normal", Some(0.0), Model.linear(Sde.ouProcess(1)) |+| Model.seasonal(24, 3, Sde.ouProcess(6)),
      Parameters(Some(0.0), ou1) |+| Parameters(None, ou6),
      s"""[{"f": "first", "sde": $ou1J}, {"f": "seasonal", "period": 24, "harmonics": 3, "sde": $ou6J}]""", 7, Vector(1.3, -0.2, 2.1, 0.4)))

  private def filterCase(c: Case, n: Int, seed: Int): String = {
    val mod = c.unparam.run(c.params).get                                     // model/Model.scala:110-136
    val filter = Filter(mod, Resampling.systematicResampling)                 // model/ParticleFilter.scala:233-246
    val d = c.d
    def replayNormals(s: Int): Vector[Vector[Double]] = { val g = new MersenneTwister(s); Vector.fill(n)(Vector.fill(d)(g.nextGaussian())) }

    Rand.generator.setSeed(seed)
    val s0 = filter.initialiseState(n, 0.0)                                   // :105-108
    val z0 = replayNormals(seed)
    var s = s0
    val steps = c.ys.zipWithIndex.map { case (y, i) =>
      val t = 0.1 * i                                                        // the first datum sits at t0: dt = 0 (:138)
      val hasObs = i != 2
      val datum: Data = TimedObservation(t, if (hasObs) Some(y) else None)
      val zseed = 100 * seed + i
      val useed = 7000 + 100 * seed + i
      // (1) the reference's own step
      Rand.generator.setSeed(zseed); scala.util.Random.setSeed(useed)
      val ref = filter.stepFilter(s, datum)                                   // :116-132
      // (2) the same step stage by stage with the reference's own functions and the same seeds
      Rand.generator.setSeed(zseed)
      val dt = t - s.t
      val x1 = s.particles map (x => filter.stepFunction(dt)(x).draw)         // :118
      val z = replayNormals(zseed)
      val fields = if (!hasObs) {
        require(x1.map(flat) == ref.particles.map(flat), s"${c.name} step $i: stage-by-stage propagate differs from stepFilter")
        ""
      } else {
        val w = x1 map (x => filter.dataLikelihood(filter.f(x, t), y))        // :123
        val mx = w.max                                                        // :124
        val w1 = w map (a => math.exp(a - mx))                                // :125
        scala.util.Random.setSeed(useed)
        val anc = Resampling.systematicResampling(Vector.range(0, n), w1)     // :126 on indices: the ancestors themselves
        val u = new scala.util.Random(useed).nextDouble()                     // model/Resampling.scala:66
        require(anc.map(x1).map(flat) == ref.particles.map(flat), s"${c.name} step $i: stage-by-stage resampling differs from stepFilter")
        val ll = s.ll + mx + math.log(ParticleFilter.mean(w1))                // :127
        require(ll == ref.ll && ParticleFilter.effectiveSampleSize(w1) == ref.ess, s"${c.name} step $i: ll / ess differ from stepFilter")
        s""", "logw": ${arr(w)}, "max": ${num(mx)}, "w1": ${arr(w1)}, "ll_incr": ${num(ref.ll - s.ll)}, "u_sys": ${num(u)}, "anc_systematic": ${iarr(anc)}"""
      }
      val out = s"""{"t": ${num(t)}, "has_obs": $hasObs, "y": ${num(y)}, "z": ${arr2(z)}, "x_prop": ${arr2(x1.map(flat))}, "ll": ${num(ref.ll)}, "ess": ${ref.ess}$fields}"""
      s = ref
      out
    }
    s"""{"model": {"name": "${c.name}", "obs": "${c.obs}", "scale": ${c.scale.map(num).getOrElse("null")}, "leaves": ${c.leavesJson}}, "N": $n, "d": $d, "euler": false, "z0": ${arr2(z0)}, "x0": ${arr2(s0.particles.map(flat))}, "t0": 0.0, "steps": ${steps.mkString("[", ", ", "]")}}"""
  }

  // ---- resampling alone: weight vectors with ties, zero runs and vanishing weights (the TreeMap quirks) --------------
  private def resamplingCase(name: String, w: Vector[Double], seed: Int): String = {
    val items = Vector.range(0, w.size)
    scala.util.Random.setSeed(seed)
    val sys = Resampling.systematicResampling(items, w)                       // model/Resampling.scala:63-72
    val u = new scala.util.Random(seed).nextDouble()
    scala.util.Random.setSeed(seed + 1)
    val strat = Resampling.stratifiedResampling(items, w)                     // :78-86
    val r2 = new scala.util.Random(seed + 1)
    val us = Vector.fill(w.size)(r2.nextDouble())
    s"""{"name": "$name", "w": ${arr(w)}, "u": ${num(u)}, "us": ${arr(us)}, "anc_systematic": ${iarr(sys)}, "anc_stratified": ${iarr(strat)}}"""
  }

  def main(args: Array[String]): Unit = {
    val dir = if (args.nonEmpty) args(0) else "."
    val filt = cases.zipWithIndex.map { case (c, i) => filterCase(c, 16, 11 + i) }
    new PrintWriter(s"$dir/filter_steps.ref.json") { write(filt.mkString("[", ",\n", "]\n")); close() }
    val ws: List[(String, Vector[Double])] = List(
      "unit" -> Vector.fill(9)(1.0),
      "geometric" -> Vector.tabulate(12)(i => math.pow(0.5, i)),
      "one_heavy_then_vanishing" -> (Vector(1.0) ++ Vector.fill(7)(1e-30)),
      "vanishing_then_heavy" -> (Vector.fill(5)(1e-25) ++ Vector(1.0, 1e-25, 1e-25)),
      "zeros_inside" -> Vector(0.3, 0.0, 0.0, 0.5, 0.0, 0.2),
      "ties" -> Vector(0.25, 0.25, 0.25, 0.25),
      "unnormalised" -> Vector(3.0, 1.0, 4.0, 1.0, 5.0, 9.0, 2.0, 6.0))
    val res = ws.zipWithIndex.map { case ((nm, w), i) => resamplingCase(nm, w, 100 + 2 * i) }
    new PrintWriter(s"$dir/resampling.ref.json") { write(res.mkString("[", ",\n", "]\n")); close() }
    println(s"wrote $dir/filter_steps.ref.json and $dir/resampling.ref.json")
  }
}
