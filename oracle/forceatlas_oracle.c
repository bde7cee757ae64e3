/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
 *
 * A plain-C, single-threaded CPU restatement of the ForceAtlas hot path of
 * LLNL/graph-embed, used as the parity checker for the CUDA implementation.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  Nothing under
 * graph-embed_b200/ may.
 *
 * Pinning: the reference holds no golden vectors for this path (its only test
 * file, examples/run-tests.cpp, covers the partitioner).  The oracle is instead
 * pinned against the reference itself: oracle/Makefile compiles the reference's
 * unmodified sources into oracle/_ref/, tests/test_oracle_vs_ref.py checks
 * BITWISE equality of positions (flat: k = 1..100 iterations; multilevel: 100
 * iterations + prolongation; whole embed() incl. the 100 000-iteration coarsest
 * solve), and tests/golden/ holds vectors minted from that compiled reference
 * (tests/golden/make_golden.py) for boxes where /root/reference is absent.
 * Operation order below is kept identical to the reference on purpose; build
 * with -ffp-contract=off (see Makefile) or bitwise parity is lost.
 *
 * All coordinates are AoS row-major n x dim, like the reference's
 * std::vector<std::vector<double>>.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int iterations;
  double ks, ksmax, repel, attract, gravity, delta, tolerate;
  int useWeights, linlog, nohubs, normalize;
} fa_params;

static const double EPSILON = 0.00001; /* include/forceatlas.hpp:110, :337 */

/* include/forceatlas.hpp:66-68 */
static double fa_abs(double v) { return (v < 0) ? -v : v; }

/* include/forceatlas.hpp:70-78  (d = v2 - v1, k ascending, then sqrt) */
static double fa_distance(const double* v1, const double* v2, int dim) {
  double sum = 0.0;
  for (int k = 0; k < dim; k++) {
    double d = v2[k] - v1[k];
    sum += d * d;
  }
  return sqrt(sum);
}

/* include/forceatlas.hpp:80-87 */
static double fa_magnitude(const double* v, int dim) {
  double sum = 0.0;
  for (int k = 0; k < dim; k++) {
    double d = v[k];
    sum += d * d;
  }
  return sqrt(sum);
}

/* include/forceatlas.hpp:176-196 (flat) == :424-444 (multilevel): attraction magnitude */
static double fa_attraction(double dis_ij, double a_weight, double deg_ip1, const fa_params* p) {
  double fa_ij = dis_ij;
  if (p->linlog) fa_ij = log(1 + fa_ij);
  double a_ij = p->useWeights ? a_weight : 1.0;
  if (p->delta == 1.0) {
    fa_ij = fa_ij * a_ij;
  } else if (p->delta != 0.0) {
    fa_ij = (a_ij < 0 ? -1 : 1) * pow(fa_abs(a_ij), p->delta) * fa_ij;
  }
  if (p->nohubs) fa_ij = fa_ij / deg_ip1;
  return p->attract * fa_ij;
}

/* include/forceatlas.hpp:127-140: weighted row sum (diagonal included) or row length */
void oracle_flat_degree(int n, const int* I, const double* D, int useWeights, double* deg) {
  for (int i = 0; i < n; i++) {
    if (useWeights) {
      double sum = 0.0;
      for (int k = I[i]; k < I[i + 1]; k++) sum += D[k];
      deg[i] = sum;
    } else {
      deg[i] = 1.0 * (I[i + 1] - I[i]);
    }
  }
}

/*
 * include/forceatlas.hpp:148-212: forces of one flat iteration for rows
 * [row_begin, row_end) from positions `coords`.  forces is n x dim (only the
 * requested rows are written).  If fscale != NULL, fscale[i] receives the sum of
 * the Euclidean norms of every individual term added into row i -- the
 * conditioning scale against which tolerance tests are stated.
 */
void oracle_flat_forces(int n, const int* I, const int* J, const double* D, int dim,
                        const double* coords, const fa_params* p, int row_begin, int row_end,
                        double* forces, double* fscale) {
  double* deg = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  oracle_flat_degree(n, I, D, p->useWeights, deg);
  double force_i[8];
  for (int i = row_begin; i < row_end; i++) {
    const double* xi = coords + (size_t)i * dim;
    for (int k = 0; k < dim; k++) force_i[k] = 0.0;
    double scale = 0.0;
    double deg_ip1 = deg[i] + 1;
    for (int j = 0; j < n; j++) { /* :151-167 */
      if (i != j) {
        const double* xj = coords + (size_t)j * dim;
        double deg_jp1 = deg[j] + 1;
        double dis_ij = fa_distance(xi, xj, dim);
        if (dis_ij < EPSILON) dis_ij = EPSILON;
        double Fr_ij = deg_ip1 * deg_jp1 * p->repel / (dis_ij * dis_ij);
        double t2 = 0.0;
        for (int k = 0; k < dim; k++) {
          double direction = -(xj[k] - xi[k]) / dis_ij;
          double Fr_sum = direction * Fr_ij;
          force_i[k] += Fr_sum;
          t2 += Fr_sum * Fr_sum;
        }
        scale += sqrt(t2);
      }
    }
    for (int k2 = I[i]; k2 < I[i + 1]; k2++) { /* :169-203 (diagonal NOT skipped) */
      int j = J[k2];
      const double* xj = coords + (size_t)j * dim;
      double dis_ij = fa_distance(xi, xj, dim);
      if (dis_ij < EPSILON) dis_ij = EPSILON;
      double Fa_ij = fa_attraction(dis_ij, D[k2], deg_ip1, p);
      double t2 = 0.0;
      for (int k = 0; k < dim; k++) {
        double direction = (xj[k] - xi[k]) / dis_ij;
        double Fa_sum = direction * Fa_ij;
        force_i[k] += Fa_sum;
        t2 += Fa_sum * Fa_sum;
      }
      scale += sqrt(t2);
    }
    double mag = fa_magnitude(xi, dim); /* :205  (unclamped) */
    double g2 = 0.0;
    for (int k = 0; k < dim; k++) {
      double Far_ki = force_i[k];
      double uv2_ki = -xi[k] / mag;
      double Fg_ki = uv2_ki * p->gravity * deg_ip1;
      forces[(size_t)i * dim + k] = Far_ki + Fg_ki;
      g2 += Fg_ki * Fg_ki;
    }
    if (fscale) fscale[i] = scale + sqrt(g2);
  }
  free(deg);
}

/* include/forceatlas.hpp:214-261: swing, (dead) global speed, per-vertex speed, move */
static void fa_step(int n, int dim, double* coords, const double* forces, const double* forces_prev,
                    const fa_params* p, int clamp_swing, const int* map) {
  double globalSpeed = p->tolerate * 1.0 / 1.0; /* :228, :242, :244 */
  for (int i = 0; i < n; i++) {
    const double* f = forces + (size_t)i * dim;
    double swing = fa_distance(f, forces_prev + (size_t)i * dim, dim);
    if (clamp_swing && swing < EPSILON) swing = EPSILON; /* :484-486 multilevel only */
    double totalF_i = fa_magnitude(f, dim);
    double speed_i = p->ks * globalSpeed / (1 + globalSpeed * sqrt(swing));
    double speedConstraint_i = p->ksmax / totalF_i;
    if (speed_i > speedConstraint_i) speed_i = speedConstraint_i;
    double* x = coords + (size_t)(map ? map[i] : i) * dim;
    for (int k = 0; k < dim; k++) {
      double displacement_ik = f[k] * speed_i;
      x[k] = displacement_ik + x[k];
    }
  }
}

/*
 * include/forceatlas.hpp:89-305 with caller-supplied initial coordinates.
 * coords: n x dim in/out.  forces_last (optional, n x dim): forces computed in
 * the final iteration (what forces_prev holds on exit).
 */
void oracle_flat_run(int n, const int* I, const int* J, const double* D, int dim, double* coords,
                     const fa_params* p, double* forces_last) {
  size_t nd = (size_t)(n > 0 ? n : 1) * dim;
  double* forces = (double*)calloc(nd, sizeof(double));
  double* forces_prev = (double*)calloc(nd, sizeof(double));
  for (int iter = 0; iter < p->iterations; iter++) {
    oracle_flat_forces(n, I, J, D, dim, coords, p, 0, n, forces, NULL);
    fa_step(n, dim, coords, forces, forces_prev, p, 0, NULL);
    memcpy(forces_prev, forces, sizeof(double) * nd); /* :263 */
  }
  if (p->normalize) { /* :272-303 */
    double avg[8] = {0};
    for (int i = 0; i < n; i++)
      for (int k = 0; k < dim; k++) avg[k] = avg[k] + coords[(size_t)i * dim + k];
    for (int k = 0; k < dim; k++) avg[k] = avg[k] / n;
    for (int i = 0; i < n; i++)
      for (int k = 0; k < dim; k++) coords[(size_t)i * dim + k] -= avg[k];
    double max_length = 0.0;
    for (int i = 0; i < n; i++) {
      double length = fa_magnitude(coords + (size_t)i * dim, dim);
      if (max_length < length) max_length = length;
    }
    for (int i = 0; i < n; i++)
      for (int k = 0; k < dim; k++)
        coords[(size_t)i * dim + k] = coords[(size_t)i * dim + k] / max_length;
  }
  if (forces_last) memcpy(forces_last, forces_prev, sizeof(double) * nd);
  free(forces);
  free(forces_prev);
}

/*
 * include/forceatlas.hpp:391-475: forces of one multilevel iteration for the
 * members v[0..s) of aggregate a, from global positions `coords`.
 * forces: s x dim (local order).  Reproduces quirk Q1 (global j compared with
 * local i at :417).
 */
static void ml_forces(int a, int s, const int* v, const double* deg, const int* I, const int* J,
                      const double* D, const int* v_A, const double* coords_A, int dim,
                      const double* coords, const fa_params* p, double* forces, double* fscale) {
  double force_i[8];
  for (int i = 0; i < s; i++) {
    const double* xi = coords + (size_t)v[i] * dim;
    for (int k = 0; k < dim; k++) force_i[k] = 0.0;
    double scale = 0.0;
    double deg_ip1 = deg[i] + 1;
    for (int j = 0; j < s; j++) { /* :394-410 */
      if (i != j) {
        const double* xj = coords + (size_t)v[j] * dim;
        double deg_jp1 = deg[j] + 1;
        double dis_ij = fa_distance(xi, xj, dim);
        if (dis_ij < EPSILON) dis_ij = EPSILON;
        double Fr_ij = deg_ip1 * deg_jp1 * p->repel / (dis_ij * dis_ij);
        double t2 = 0.0;
        for (int k = 0; k < dim; k++) {
          double direction = -(xj[k] - xi[k]) / dis_ij;
          double Fr_sum = direction * Fr_ij;
          force_i[k] += Fr_sum;
          t2 += Fr_sum * Fr_sum;
        }
        scale += sqrt(t2);
      }
    }
    double mag = fa_magnitude(xi, dim); /* :411-414 (clamped here) */
    if (mag < EPSILON) mag = EPSILON;
    for (int k2 = I[v[i]]; k2 < I[v[i] + 1]; k2++) { /* :415-467 */
      int j = J[k2];
      double t2 = 0.0;
      if (v_A[j] == a && j != i) { /* internal; Q1: global j vs local i */
        const double* xj = coords + (size_t)j * dim;
        double dis_ij = fa_distance(xi, xj, dim);
        if (dis_ij < EPSILON) dis_ij = EPSILON;
        double Fa_ij = fa_attraction(dis_ij, D[k2], deg_ip1, p);
        for (int k = 0; k < dim; k++) {
          double direction = (xj[k] - xi[k]) / dis_ij;
          double Fa_sum = direction * Fa_ij;
          force_i[k] += Fa_sum;
          t2 += Fa_sum * Fa_sum;
        }
      } else { /* external pull toward the neighbouring aggregate's centre */
        const double* ca = coords_A + (size_t)a * dim;
        const double* cb = coords_A + (size_t)v_A[j] * dim;
        double pull = 100.0;
        double dis_ij = fa_distance(ca, cb, dim);
        if (dis_ij < EPSILON) dis_ij = EPSILON;
        double fao_ij = 1.0;
        double Fao_ij = pull * fao_ij;
        for (int k = 0; k < dim; k++) {
          double direction = (cb[k] - ca[k]) / dis_ij;
          double Fao_sum = direction * Fao_ij / mag;
          force_i[k] += Fao_sum;
          t2 += Fao_sum * Fao_sum;
        }
      }
      scale += sqrt(t2);
    }
    double g2 = 0.0;
    for (int k = 0; k < dim; k++) { /* :469-474 */
      double Far_ki = force_i[k];
      double uv2_ki = -xi[k] / mag;
      double Fg_ki = uv2_ki * p->gravity * deg_ip1;
      forces[(size_t)i * dim + k] = Far_ki + Fg_ki;
      g2 += Fg_ki * Fg_ki;
    }
    if (fscale) fscale[v[i]] = scale + sqrt(g2);
  }
}

/*
 * include/forceatlas.hpp:314-574 with the random initial local coordinates
 * supplied by the caller: init is n x dim indexed by GLOBAL vertex id (the
 * reference draws them aggregate-major, member-major, k-minor at :356-360).
 *
 * coords_out: n x dim.  If forces_at != NULL (n x dim, global order) it receives
 * the forces computed in iteration number `forces_iter` (0-based) and
 * fscale_at (n) the matching conditioning scales.
 */
void oracle_multilevel_run_range(int n, const int* I, const int* J, const double* D, int m,
                                 const int* PI, const int* PJ, const int* v_A,
                                 const double* coords_A, const double* r_A, int dim,
                                 const double* init, const fa_params* p, double* coords_out,
                                 int forces_iter, double* forces_at, double* fscale_at,
                                 int a_begin, int a_end) {
  double* coords = coords_out;
  (void)n;
  (void)m;
  /* the aggregates are independent (:340-341): a sub-range leaves the other rows untouched */
  for (int a = a_begin; a < a_end; a++) {
    const int* v = PJ + PI[a];
    int s = PI[a + 1] - PI[a];
    for (int i = 0; i < s; i++)
      for (int k = 0; k < dim; k++)
        coords[(size_t)v[i] * dim + k] = init[(size_t)v[i] * dim + k]; /* :356-360 */

    double* deg = (double*)malloc(sizeof(double) * (size_t)(s > 0 ? s : 1));
    for (int i = 0; i < s; i++) { /* :362-383 intra-aggregate degree (self-loops included) */
      double sum = 0.0;
      for (int k = I[v[i]]; k < I[v[i] + 1]; k++)
        if (v_A[J[k]] == a) sum += p->useWeights ? D[k] : 1.0;
      deg[i] = sum;
    }
    size_t sd = (size_t)(s > 0 ? s : 1) * dim;
    double* forces = (double*)calloc(sd, sizeof(double));
    double* forces_prev = (double*)calloc(sd, sizeof(double));
    for (int iter = 0; iter < p->iterations; iter++) { /* :390-538 */
      int want = (forces_at != NULL && iter == forces_iter);
      ml_forces(a, s, v, deg, I, J, D, v_A, coords_A, dim, coords, p, forces,
                want ? fscale_at : NULL);
      if (want)
        for (int i = 0; i < s; i++)
          for (int k = 0; k < dim; k++)
            forces_at[(size_t)v[i] * dim + k] = forces[(size_t)i * dim + k];
      fa_step(s, dim, coords, forces, forces_prev, p, 1, v);
      memcpy(forces_prev, forces, sizeof(double) * sd);
    }
    { /* :539-570 centre, max-normalise, prolongate into the parent ball */
      double avg[8] = {0};
      for (int i = 0; i < s; i++)
        for (int k = 0; k < dim; k++) avg[k] = avg[k] + coords[(size_t)v[i] * dim + k];
      for (int k = 0; k < dim; k++) avg[k] = avg[k] / s;
      for (int i = 0; i < s; i++)
        for (int k = 0; k < dim; k++) coords[(size_t)v[i] * dim + k] -= avg[k];
      double max = 0.0;
      for (int i = 0; i < s; i++) {
        double sum = fa_magnitude(coords + (size_t)v[i] * dim, dim);
        if (sum > max) max = sum;
      }
      if (max < EPSILON) max = EPSILON;
      for (int i = 0; i < s; i++)
        for (int k = 0; k < dim; k++)
          coords[(size_t)v[i] * dim + k] =
              coords_A[(size_t)a * dim + k] + r_A[a] * (coords[(size_t)v[i] * dim + k] / max);
    }
    free(deg);
    free(forces);
    free(forces_prev);
  }
}

void oracle_multilevel_run(int n, const int* I, const int* J, const double* D, int m,
                           const int* PI, const int* PJ, const int* v_A, const double* coords_A,
                           const double* r_A, int dim, const double* init, const fa_params* p,
                           double* coords_out, int forces_iter, double* forces_at,
                           double* fscale_at) {
  oracle_multilevel_run_range(n, I, J, D, m, PI, PJ, v_A, coords_A, r_A, dim, init, p, coords_out,
                              forces_iter, forces_at, fscale_at, 0, m);
}

/* ------------------------------------------------------------------------- */
/* Level driver: src/embed.cpp:576-796                                        */
/* ------------------------------------------------------------------------- */

typedef struct {
  double t;
  int i, j;
} fa_event;

/* std::sort on std::tuple<double,int,int> is lexicographic ascending */
static int fa_event_cmp(const void* pa, const void* pb) {
  const fa_event* a = (const fa_event*)pa;
  const fa_event* b = (const fa_event*)pb;
  if (a->t < b->t) return -1;
  if (a->t > b->t) return 1;
  if (a->i != b->i) return a->i < b->i ? -1 : 1;
  if (a->j != b->j) return a->j < b->j ? -1 : 1;
  return 0;
}

/*
 * The "ball growing" event loop shared by src/embed.cpp:636-678 (base case) and
 * :713-755 (general case): pop the latest-sorted event, freeze the live
 * endpoint(s) at that radius, push back the events that touch them, re-sort.
 */
static void fa_grow_balls(fa_event* times, size_t count_events, double* r_A, int m) {
  qsort(times, count_events, sizeof(fa_event), fa_event_cmp);
  int count = 0;
  size_t sz = count_events;
  while (count < m && sz != 0) {
    fa_event e = times[sz - 1];
    double time_ij = e.t;
    int i = e.i, j = e.j;
    double distance = -time_ij;
    sz--;
    int live_i = (r_A[i] <= 0.0), live_j = (r_A[j] <= 0.0);
    if (live_i && !live_j) {
      r_A[i] = distance;
      for (size_t a = 0; a < sz; a++)
        if (times[a].i == i || times[a].j == i) times[a].t = -(2 * (-times[a].t) - (-time_ij));
      qsort(times, sz, sizeof(fa_event), fa_event_cmp);
      count++;
    } else if (!live_i && live_j) {
      r_A[j] = distance;
      for (size_t a = 0; a < sz; a++)
        if (times[a].i == j || times[a].j == j) times[a].t = -(2 * (-times[a].t) - (-time_ij));
      qsort(times, sz, sizeof(fa_event), fa_event_cmp);
      count++;
    } else if (live_i && live_j) {
      r_A[i] = distance;
      r_A[j] = distance;
      for (size_t a = 0; a < sz; a++)
        if (times[a].i == i || times[a].j == i || times[a].i == j || times[a].j == j)
          times[a].t = -(2 * (-times[a].t) - (-time_ij));
      qsort(times, sz, sizeof(fa_event), fa_event_cmp);
      count += 2;
    }
  }
}

/*
 * src/embed.cpp:615-778: radii r_A of the m vertices of level+1 (whose
 * coordinates coords_A were just computed) and, in the general case, the
 * rescale of coords_A / r_A into the grand-parent balls.
 *
 *  base case (mc == 0, r_Ac == NULL): all pairs, :616-679
 *  general case: Ac = CSR of A_{level+1} (m rows); PcI/PcJ = P_T of level+1
 *    (mc rows over m columns); coords_Ac (mc x dim), r_Ac (mc): :680-777
 */
void oracle_radii(int m, int dim, double* coords_A, double* r_A, const int* AcI, const int* AcJ,
                  int mc, const int* PcI, const int* PcJ, const double* coords_Ac,
                  const double* r_Ac) {
  for (int i = 0; i < m; i++) r_A[i] = 0.0;
  if (mc == 0) {
    size_t cap = (size_t)m * (size_t)(m > 0 ? m - 1 : 0) / 2;
    fa_event* times = (fa_event*)malloc(sizeof(fa_event) * (cap > 0 ? cap : 1));
    size_t cnt = 0;
    for (int i = 0; i < m; i++)
      for (int j = i + 1; j < m; j++) {
        double distance_ij =
            fa_distance(coords_A + (size_t)i * dim, coords_A + (size_t)j * dim, dim);
        times[cnt].t = -distance_ij / 2;
        times[cnt].i = i;
        times[cnt].j = j;
        cnt++;
      }
    fa_grow_balls(times, cnt, r_A, m);
    free(times);
    return;
  }
  int* vertex_Ac = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  for (int b = 0; b < mc; b++)
    for (int c = PcI[b]; c < PcI[b + 1]; c++) vertex_Ac[PcJ[c]] = b; /* :684 */
  for (int b = 0; b < mc; b++) {                                     /* :686-756 */
    int s = PcI[b + 1] - PcI[b];
    if (s == 1) {
      r_A[PcJ[PcI[b]]] = r_Ac[b];
      continue;
    }
    size_t cap = 0;
    for (int i = 0; i < s; i++) {
      int a = PcJ[PcI[b] + i];
      cap += (size_t)(AcI[a + 1] - AcI[a]);
    }
    fa_event* times = (fa_event*)malloc(sizeof(fa_event) * (cap > 0 ? cap : 1));
    size_t cnt = 0;
    for (int i = 0; i < s; i++) {
      int a = PcJ[PcI[b] + i];
      for (int kk = AcI[a]; kk < AcI[a + 1]; kk++) {
        int j = AcJ[kk];
        if (a < j && vertex_Ac[j] == vertex_Ac[a]) {
          double distance_ij =
              fa_distance(coords_A + (size_t)a * dim, coords_A + (size_t)j * dim, dim);
          times[cnt].t = -distance_ij / 2;
          times[cnt].i = a;
          times[cnt].j = j;
          cnt++;
        }
      }
    }
    fa_grow_balls(times, cnt, r_A, m);
    free(times);
  }
  for (int b = 0; b < mc; b++) { /* :757-777 */
    const double* cb = coords_Ac + (size_t)b * dim;
    double alpha = 0.0;
    for (int k2 = PcI[b]; k2 < PcI[b + 1]; k2++) {
      int a = PcJ[k2];
      double dis = fa_distance(cb, coords_A + (size_t)a * dim, dim) + r_A[a];
      if (dis > alpha) alpha = dis;
    }
    double epsilon = 0.000001;
    if (alpha < epsilon) alpha = epsilon;
    for (int k2 = PcI[b]; k2 < PcI[b + 1]; k2++) {
      int a = PcJ[k2];
      for (int k = 0; k < dim; k++)
        coords_A[(size_t)a * dim + k] =
            cb[k] + (r_Ac[b] / alpha) * (coords_A[(size_t)a * dim + k] - cb[k]);
      r_A[a] = (r_Ac[b] / alpha) * r_A[a];
    }
  }
  free(vertex_Ac);
}

/* ------------------------------------------------------------------------- */
/* std::mt19937 + libstdc++ uniform_real_distribution<double>(-1,1)           */
/* (the reference's generator, include/forceatlas.hpp:104-108, :332-336)      */
/* ------------------------------------------------------------------------- */

typedef struct {
  uint32_t mt[624];
  int idx;
} mt19937_state;

static void mt_seed(mt19937_state* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
}

static uint32_t mt_next(mt19937_state* s) {
  if (s->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
      s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* generate_canonical<double,53>: two 32-bit draws, low word first; then a + (b-a)*u */
void oracle_mt19937_uniform(uint32_t seed, size_t count, double* out) {
  mt19937_state s;
  mt_seed(&s, seed);
  for (size_t c = 0; c < count; c++) {
    double lo = (double)mt_next(&s);
    double hi = (double)mt_next(&s);
    double sum = lo + hi * 4294967296.0;
    double u = sum / 18446744073709551616.0;
    if (u >= 1.0) u = nextafter(1.0, 0.0);
    out[c] = u * (1.0 - (-1.0)) + (-1.0);
  }
}

/*
 * src/embed.cpp:561-796 `embed`, with every random stream replaced by
 * mt19937(seed) (each generator the reference constructs starts from the same
 * seed; see oracle/ref_prelude.hpp).  As/Ps are arrays of CSR pointers:
 * AI[l],AJ[l],AD[l] with An[l] rows for l = 0..L ; PI[l],PJ[l] with Pm[l] rows
 * for l = 0..L-1.  coords_out: An[0] x dim.  Returns 0.
 */
static double* embed_level(int L, const int* An, const int* const* AI, const int* const* AJ,
                           const double* const* AD, const int* Pm, const int* const* PI,
                           const int* const* PJ, int dim, int level, uint32_t seed,
                           int coarse_iterations, int level_iterations, double** r_A_out,
                           double** coords_A_out) {
  fa_params p = {0, 0.1, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1, 0, 0, 0};
  int n = An[level];
  double* coords = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * dim);
  if (level == L) { /* :582-587 -> forceatlas.hpp:307-312 (defaults, 100 000 iterations) */
    *r_A_out = NULL;
    *coords_A_out = NULL;
    oracle_mt19937_uniform(seed, (size_t)n * dim, coords);
    p.iterations = coarse_iterations;
    oracle_flat_run(n, AI[level], AJ[level], AD[level], dim, coords, &p, NULL);
    return coords;
  }
  double *r_Ac = NULL, *coords_Ac = NULL;
  double* coords_A = embed_level(L, An, AI, AJ, AD, Pm, PI, PJ, dim, level + 1, seed,
                                 coarse_iterations, level_iterations, &r_Ac, &coords_Ac);
  int m = An[level + 1];
  double* r_A = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  if (r_Ac == NULL)
    oracle_radii(m, dim, coords_A, r_A, NULL, NULL, 0, NULL, NULL, NULL, NULL);
  else
    oracle_radii(m, dim, coords_A, r_A, AI[level + 1], AJ[level + 1], Pm[level + 1],
                 PI[level + 1], PJ[level + 1], coords_Ac, r_Ac);
  int* v_A = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1)); /* :605 */
  for (int a = 0; a < Pm[level]; a++)
    for (int c = PI[level][a]; c < PI[level][a + 1]; c++) v_A[PJ[level][c]] = a;
  /* draw order: aggregate-major, member-major, k-minor (forceatlas.hpp:341,356-358) */
  double* stream = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * dim);
  double* init = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * dim);
  oracle_mt19937_uniform(seed, (size_t)n * dim, stream);
  size_t pos = 0;
  for (int a = 0; a < Pm[level]; a++)
    for (int c = PI[level][a]; c < PI[level][a + 1]; c++)
      for (int k = 0; k < dim; k++) init[(size_t)PJ[level][c] * dim + k] = stream[pos++];
  p.iterations = level_iterations; /* :793 */
  oracle_multilevel_run(n, AI[level], AJ[level], AD[level], Pm[level], PI[level], PJ[level], v_A,
                        coords_A, r_A, dim, init, &p, coords, 0, NULL, NULL);
  free(stream);
  free(init);
  free(v_A);
  free(r_Ac);
  free(coords_Ac);
  *r_A_out = r_A;
  *coords_A_out = coords_A;
  return coords;
}

int oracle_embed(int L, const int* An, const int* const* AI, const int* const* AJ,
                 const double* const* AD, const int* Pm, const int* const* PI,
                 const int* const* PJ, int dim, uint32_t seed, int coarse_iterations,
                 int level_iterations, double* coords_out) {
  double *r_A = NULL, *coords_A = NULL;
  double* coords = embed_level(L, An, AI, AJ, AD, Pm, PI, PJ, dim, 0, seed, coarse_iterations,
                               level_iterations, &r_A, &coords_A);
  memcpy(coords_out, coords, sizeof(double) * (size_t)An[0] * dim);
  free(coords);
  free(r_A);
  free(coords_A);
  return 0;
}


/* ---------------------------------------------------------------------------------------------
 * Galerkin coarse graph, examples/embedder.cpp:213-216 / examples/embed.cpp:95-98:
 *     As.push_back(P.Mult(As.back()).Mult(P.Transpose()))
 * for a 0/1 aggregation P (m x n, one entry per column).  linalgcpp (where Mult lives) is not in
 * the reference tree, so this restates the product itself: A_c[a][b] = sum of A[i][j] over i in a,
 * j in b, accumulated row by row (members of a in P's order, entries of row i in CSR order) into a
 * dense accumulator, emitted with ascending columns.  PARITY UNPINNED against linalgcpp's own
 * summation order; for unit-weight graphs (every BASELINE config) the sums are integers and the
 * result does not depend on the order.  out_idx / out_val need room for nnz(A) entries.
 * Returns nnz(A_c). */
long oracle_galerkin(int n, int m, const int* I, const int* J, const double* D, const int* PI,
                     const int* PJ, int* out_ptr, int* out_idx, double* out_val) {
  int* vA = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  double* acc = (double*)calloc((size_t)(m > 0 ? m : 1), sizeof(double));
  int* mark = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  int* cols = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  for (int a = 0; a < m; ++a) {
    mark[a] = -1;
    for (int c = PI[a]; c < PI[a + 1]; ++c) vA[PJ[c]] = a;
  }
  long nnz = 0;
  out_ptr[0] = 0;
  for (int a = 0; a < m; ++a) {
    int ncols = 0;
    for (int c = PI[a]; c < PI[a + 1]; ++c) {
      const int i = PJ[c];
      for (int e = I[i]; e < I[i + 1]; ++e) {
        const int b = vA[J[e]];
        if (mark[b] != a) {
          mark[b] = a;
          acc[b] = 0.0;
          cols[ncols++] = b;
        }
        acc[b] += D ? D[e] : 1.0;
      }
    }
    /* ascending columns (insertion sort: rows are short) */
    for (int x = 1; x < ncols; ++x) {
      const int v = cols[x];
      int y = x - 1;
      while (y >= 0 && cols[y] > v) {
        cols[y + 1] = cols[y];
        --y;
      }
      cols[y + 1] = v;
    }
    for (int x = 0; x < ncols; ++x) {
      out_idx[nnz] = cols[x];
      out_val[nnz] = acc[cols[x]];
      ++nnz;
    }
    out_ptr[a + 1] = (int)nnz;
  }
  free(vA);
  free(acc);
  free(mark);
  free(cols);
  return nnz;
}
