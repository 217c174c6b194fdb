// hk_scene.cuh -- the fixed HockeyEnv scene as flat tables (reference hockey_env.py:183-343,373-406).
//
// The reference rebuilds 16 Box2D bodies through SWIG on every reset (hockey_env.py:162-178,371-406);
// here the geometry is one immutable table in __constant__ memory and a reset only rewrites per-env
// state.  Fixture numbering = creation order of the colliding fixtures (decorations have
// categoryBits = maskBits = 0, hockey_env.py:246-247, and never collide):
//   0 wall top, 1 wall bottom, 2 left-top, 3 left-bottom, 4 right-top, 5 right-bottom,
//   6 goal1 sensor, 7 goal1 solid, 8 goal2 sensor, 9 goal2 solid, 10 racket1, 11 racket2, 12 puck.
// Contact pair ids (fixed candidate set) are documented in include/hockey_b200.h.
#pragma once
#include "hk_math.cuh"

namespace hk {

enum { F_G1_SENSOR = 6, F_G1_SOLID = 7, F_G2_SENSOR = 8, F_G2_SOLID = 9, F_R1 = 10, F_R2 = 11, F_PUCK = 12 };
enum { N_STATIC_FIX = 10, N_POLY = 12, N_PAIRS = 27 };
enum { B_R1 = 0, B_R2 = 1, B_PUCK = 2 };  // dynamic bodies

struct Poly {
  int count;
  float vx[8], vy[8], nx[8], ny[8];
  float cenx, ceny;  // b2PolygonShape::m_centroid
};

struct Scene {
  Poly poly[N_POLY];
  float spx[N_STATIC_FIX], spy[N_STATIC_FIX];  // static body origin of each static fixture
  AABB sfat[N_STATIC_FIX];                     // static broad-phase fat AABBs
  float mass[3], invMass[3], inertia[3], invI[3];  // racket1, racket2, puck
  float lcx[3], lcy[3];                            // b2Sweep::localCenter
  float puckRadius;
  float friction[N_PAIRS], restitution[N_PAIRS];  // b2MixFriction / b2MixRestitution per pair
  unsigned char pairFA[N_PAIRS], pairFB[N_PAIRS];  // fixture A/B of each pair (A as b2Contact::Create orders it)
  unsigned char sortedPairs[N_PAIRS];              // pair ids in (fixtureA, fixtureB) lexicographic order
};

// pair id -> fixtures
HK_HD void pairFixtures(int pid, int* fA, int* fB) {
  if (pid < 8) {
    *fA = pid < 6 ? pid : (pid == 6 ? F_G1_SOLID : F_G2_SOLID);
    *fB = F_R1;
  } else if (pid < 16) {
    int s = pid - 8;
    *fA = s < 6 ? s : (s == 6 ? F_G1_SOLID : F_G2_SOLID);
    *fB = F_R2;
  } else if (pid == 16) {
    *fA = F_R1;
    *fB = F_R2;
  } else if (pid < 23) {
    *fA = pid - 17;
    *fB = F_PUCK;
  } else if (pid == 23) {
    *fA = F_G1_SENSOR;
    *fB = F_PUCK;
  } else if (pid == 24) {
    *fA = F_G2_SENSOR;
    *fB = F_PUCK;
  } else if (pid == 25) {
    *fA = F_R1;
    *fB = F_PUCK;
  } else {
    *fA = F_R2;
    *fB = F_PUCK;
  }
}
// fixture -> dynamic body index, or -1 for static
HK_HD int fixtureBody(int f) { return f < N_STATIC_FIX ? -1 : f - F_R1; }

#define HK_PAIRS_R1 0x020100FFu      /* pairs touching racket1: 0-7, 16, 25 */
#define HK_PAIRS_R2 0x0401FF00u      /* pairs touching racket2: 8-15, 16, 26 */
#define HK_PAIRS_PUCK 0x07FE0000u    /* pairs touching the puck: 17-26 */
#define HK_PAIRS_SENSOR 0x01800000u  /* 23, 24 */
#define HK_PAIRS_TOI 0x007EFFFFu     /* static x dynamic, non-sensor: 0-15, 17-22 */

// ---- host-side scene construction: restates b2PolygonShape::Set / ComputeMass / b2CircleShape::
// ComputeMass / b2Body::ResetMassData (Box2D 2.3.0) in float32 for the vertex lists of
// hockey_env.py:31-32,188-195,209-214,228-232,307-317,327-338,373-375. ---------------------------
namespace scene_build {

struct TmpPoly {
  int n;
  V2 v[8], nrm[8];
  V2 centroid;
};

inline void setPolygon(TmpPoly* out, const V2* in, int cnt) {
  V2 ps[8];
  int tempCount = 0;
  for (int i = 0; i < cnt; ++i) {
    bool unique = true;
    for (int j = 0; j < tempCount; ++j)
      if (distanceSq(in[i], ps[j]) < 0.5f * HK_LINEAR_SLOP) unique = false;
    if (unique) ps[tempCount++] = in[i];
  }
  int n = tempCount;
  int i0 = 0;
  float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) {
      i0 = i;
      x0 = x;
    }
  }
  int hull[8];
  int m = 0, ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) {
        ie = j;
        continue;
      }
      V2 r = ps[ie] - ps[hull[m]];
      V2 v = ps[j] - ps[hull[m]];
      float c = cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && lengthSq(v) > lengthSq(r)) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  out->n = m;
  for (int i = 0; i < m; ++i) out->v[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    int i2 = i + 1 < m ? i + 1 : 0;
    V2 edge = out->v[i2] - out->v[i];
    out->nrm[i] = cross(edge, 1.0f);
    normalize(out->nrm[i]);
  }
  V2 c = mk(0.0f, 0.0f);
  float area = 0.0f;
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < m; ++i) {
    V2 p1 = mk(0.0f, 0.0f), p2 = out->v[i], p3 = i + 1 < m ? out->v[i + 1] : out->v[0];
    V2 e1 = p2 - p1, e2 = p3 - p1;
    float D = cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    c += (triangleArea * inv3) * (p1 + p2 + p3);
  }
  c *= 1.0f / area;
  out->centroid = c;
}

inline void polygonMass(const TmpPoly& p, float density, float* mass, V2* center, float* I) {
  V2 c = mk(0.0f, 0.0f);
  float area = 0.0f, inertia = 0.0f;
  V2 s = mk(0.0f, 0.0f);
  for (int i = 0; i < p.n; ++i) s += p.v[i];
  s *= 1.0f / p.n;
  const float k_inv3 = 1.0f / 3.0f;
  for (int i = 0; i < p.n; ++i) {
    V2 e1 = p.v[i] - s;
    V2 e2 = i + 1 < p.n ? p.v[i + 1] - s : p.v[0] - s;
    float D = cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    c += (triangleArea * k_inv3) * (e1 + e2);
    float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
    float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
    float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
    inertia += (0.25f * k_inv3 * D) * (intx2 + inty2);
  }
  *mass = density * area;
  c *= 1.0f / area;
  *center = c + s;
  *I = density * inertia;
  *I += *mass * (dot(*center, *center) - dot(c, c));
}

inline void storePoly(Scene* S, int f, const TmpPoly& t) {
  Poly& p = S->poly[f];
  p.count = t.n;
  for (int i = 0; i < 8; ++i) {
    int k = i < t.n ? i : 0;
    p.vx[i] = t.v[k].x;
    p.vy[i] = t.v[k].y;
    p.nx[i] = t.nrm[k].x;
    p.ny[i] = t.nrm[k].y;
  }
  p.cenx = t.centroid.x;
  p.ceny = t.centroid.y;
}

inline void staticFixture(Scene* S, int f, double px, double py, const double (*pts)[2], double sx, double sy) {
  const double SCALE = 60.0;
  V2 in[4];
  for (int i = 0; i < 4; ++i) in[i] = mk((float)(sx * pts[i][0] / SCALE), (float)(sy * pts[i][1] / SCALE));
  TmpPoly t;
  setPolygon(&t, in, 4);
  storePoly(S, f, t);
  S->spx[f] = (float)px;
  S->spy[f] = (float)py;
  // b2PolygonShape::ComputeAABB at the body transform (angle 0), then b2DynamicTree::CreateProxy
  Xf xf;
  xf.p = mk((float)px, (float)py);
  xf.q = rotOf(0.0f);
  V2 lower = mul(xf, t.v[0]), upper = lower;
  for (int i = 1; i < t.n; ++i) {
    V2 v = mul(xf, t.v[i]);
    lower = mk(fmin2(lower.x, v.x), fmin2(lower.y, v.y));
    upper = mk(fmax2(upper.x, v.x), fmax2(upper.y, v.y));
  }
  S->sfat[f].lx = (lower.x - HK_POLYGON_RADIUS) - HK_AABB_EXTENSION;
  S->sfat[f].ly = (lower.y - HK_POLYGON_RADIUS) - HK_AABB_EXTENSION;
  S->sfat[f].hx = (upper.x + HK_POLYGON_RADIUS) + HK_AABB_EXTENSION;
  S->sfat[f].hy = (upper.y + HK_POLYGON_RADIUS) + HK_AABB_EXTENSION;
}

inline void build(Scene* S) {
  const double SCALE = 60.0, W = 600 / SCALE, H = 480 / SCALE, GOAL_SIZE = 75, RACKETFACTOR = 1.2;
  const double wall[4][2] = {{-250, 10}, {-250, -10}, {250, -10}, {250, 10}};
  staticFixture(S, 0, W / 2, H - .5, wall, 1, 1);
  staticFixture(S, 1, W / 2, .5, wall, 1, 1);
  const double cp[4][2] = {{-10, (H - 1) / 2 * SCALE - GOAL_SIZE}, {10, (H - 1) / 2 * SCALE - GOAL_SIZE - 7}, {10, -5}, {-10, -5}};
  staticFixture(S, 2, W / 2 - 245 / SCALE, H - .5, cp, 1, -1);
  staticFixture(S, 3, W / 2 - 245 / SCALE, .5, cp, 1, 1);
  staticFixture(S, 4, W / 2 + 245 / SCALE, H - .5, cp, -1, -1);
  staticFixture(S, 5, W / 2 + 245 / SCALE, 0.5, cp, -1, 1);
  const double goal[4][2] = {{-10, GOAL_SIZE}, {10, GOAL_SIZE}, {10, -GOAL_SIZE}, {-10, -GOAL_SIZE}};
  staticFixture(S, F_G1_SENSOR, W / 2 - 245 / SCALE - 10 / SCALE, H / 2, goal, 1, 1);
  staticFixture(S, F_G1_SOLID, W / 2 - 245 / SCALE - 10 / SCALE, H / 2, goal, 1, 1);
  staticFixture(S, F_G2_SENSOR, W / 2 + 245 / SCALE + 10 / SCALE, H / 2, goal, 1, 1);
  staticFixture(S, F_G2_SOLID, W / 2 + 245 / SCALE + 10 / SCALE, H / 2, goal, 1, 1);
  const double RACKETPOLY[7][2] = {{-10, 20}, {+5, 20}, {+5, -20}, {-10, -20}, {-18, -10}, {-21, 0}, {-18, 10}};
  for (int k = 0; k < 2; ++k) {
    V2 in[7];
    for (int i = 0; i < 7; ++i) {
      double x = k ? -RACKETPOLY[i][0] / SCALE * RACKETFACTOR : RACKETPOLY[i][0] / SCALE * RACKETFACTOR;
      double y = RACKETPOLY[i][1] / SCALE * RACKETFACTOR;
      in[i] = mk((float)x, (float)y);
    }
    TmpPoly t;
    setPolygon(&t, in, 7);
    storePoly(S, F_R1 + k, t);
    // b2Body::ResetMassData for the single fixture
    float m, I;
    V2 center;
    polygonMass(t, (float)(200.0 / RACKETFACTOR), &m, &center, &I);
    float mass = 0.0f + m;
    V2 lc = mk(0.0f, 0.0f);
    lc += m * center;
    float inertia = 0.0f + I;
    float invMass = 1.0f / mass;
    lc *= invMass;
    inertia -= mass * dot(lc, lc);
    S->mass[k] = mass;
    S->invMass[k] = invMass;
    S->inertia[k] = inertia;
    S->invI[k] = 1.0f / inertia;
    S->lcx[k] = lc.x;
    S->lcy[k] = lc.y;
  }
  {
    float r = (float)(13 / SCALE);
    S->puckRadius = r;
    float m = 7.0f * HK_PI * r * r;
    float I = m * (0.5f * r * r + 0.0f);
    float mass = 0.0f + m;
    V2 lc = mk(0.0f, 0.0f);
    lc += m * mk(0.0f, 0.0f);
    float inertia = 0.0f + I;
    float invMass = 1.0f / mass;
    lc *= invMass;
    inertia -= mass * dot(lc, lc);
    S->mass[2] = mass;
    S->invMass[2] = invMass;
    S->inertia[2] = inertia;
    S->invI[2] = 1.0f / inertia;
    S->lcx[2] = lc.x;
    S->lcy[2] = lc.y;
  }
  // materials (hockey_env.py:191-195, 210-214, 229-232, 328-338)
  float fr[13], re[13];
  for (int f = 0; f < 10; ++f) {
    fr[f] = 0.1f;
    re[f] = 0.0f;
  }
  fr[F_R1] = fr[F_R2] = 1.0f;
  re[F_R1] = re[F_R2] = 0.0f;
  fr[F_PUCK] = 0.1f;
  re[F_PUCK] = 0.95f;
  int keys[N_PAIRS];
  for (int pid = 0; pid < N_PAIRS; ++pid) {
    int fA, fB;
    pairFixtures(pid, &fA, &fB);
    S->pairFA[pid] = (unsigned char)fA;
    S->pairFB[pid] = (unsigned char)fB;
    S->friction[pid] = sqrtf(fr[fA] * fr[fB]);
    S->restitution[pid] = re[fA] > re[fB] ? re[fA] : re[fB];
    keys[pid] = fA * 16 + fB;
    S->sortedPairs[pid] = (unsigned char)pid;
  }
  for (int i = 1; i < N_PAIRS; ++i) {  // insertion sort by (fixtureA, fixtureB)
    unsigned char p = S->sortedPairs[i];
    int j = i - 1;
    while (j >= 0 && keys[S->sortedPairs[j]] > keys[p]) {
      S->sortedPairs[j + 1] = S->sortedPairs[j];
      --j;
    }
    S->sortedPairs[j + 1] = p;
  }
}

}  // namespace scene_build

}  // namespace hk
