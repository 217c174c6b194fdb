// oracle/b2mini.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the parts of the Box2D 2.3.0-era rigid-body engine that the reference
// `hockey/hockey_env.py` exercises through `box2d-py` (an un-vendored third-party dependency,
// unpinned in reference `setup.py:11`; SURVEY.md section 8c names it; SURVEY.md A.8 shows the wrapped
// engine uses the 2.3.0 clamp-form damping).  The engine source is NOT in /root/reference, so this
// file restates its *published algorithm* (b2World::Step = Collide -> Solve islands -> SolveTOI),
// anchored on the reference's own call site `hockey_env.py:682` (world.Step(1/50, 180, 60)).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// compile, link or call this.  The product (hockey_env_b200/csrc) never includes it.
//
// Parity status: pinned against the reference's only exact artefact (the TRAIN_DEFENSE reward
// trace of Hockey-Env.ipynb cell 20, tests/golden/notebook_fixtures.json) and the notebook's
// statistical outputs; polygon-polygon contacts, the 2-point block solver and TOI are "parity
// unpinned" by the reference itself (it has no tests) -- see DESIGN.md.
//
// All arithmetic is scalar IEEE float32 in Box2D's own operation order; compile with
// -ffp-contract=off (x86-64 pybox2d wheels have no FMA contraction).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace b2mini {

// ---- b2Settings.h ----------------------------------------------------------------------------
static const float kEps = FLT_EPSILON;
static const float kMaxFloat = FLT_MAX;
static const float kPi = 3.14159265359f;
static const float kLinearSlop = 0.005f;
static const float kAngularSlop = (2.0f / 180.0f * kPi);
static const float kPolygonRadius = (2.0f * kLinearSlop);
static const int kMaxPolygonVertices = 16;  // pybox2d builds with 16 (upstream default 8)
static const float kVelocityThreshold = 1.0f;
static const float kMaxLinearCorrection = 0.2f;
static const float kMaxTranslation = 2.0f;
static const float kMaxTranslationSquared = (kMaxTranslation * kMaxTranslation);
static const float kMaxRotation = (0.5f * kPi);
static const float kMaxRotationSquared = (kMaxRotation * kMaxRotation);
static const float kBaumgarte = 0.2f;
static const float kToiBaumgarte = 0.75f;
static const float kTimeToSleep = 0.5f;
static const float kLinearSleepTolerance = 0.01f;
static const float kAngularSleepTolerance = (2.0f / 180.0f * kPi);
static const float kAabbExtension = 0.1f;
static const float kAabbMultiplier = 2.0f;
static const int kMaxSubSteps = 8;
static const int kMaxTOIContacts = 32;

// ---- trig ------------------------------------------------------------------------------------
// b2Rot::Set calls sinf/cosf.  Mode 1 uses libm.  Mode 0 (default) evaluates sin/cos in double
// with a fixed polynomial and rounds to float: the result is the correctly rounded float in all
// but ~1e-8 of arguments (tests/test_oracle_trig.py measures this against libm), and -- unlike
// libm -- is bit-reproducible on the GPU, which lets the CUDA path be compared bit-for-bit.  (Measured: the
// polynomial gives the correctly rounded value; glibc sinf/cosf is 1 ulp off it for ~1 % of arguments.)
extern int g_trig_mode;
// Box2D 2.3.0's b2Sweep::Advance computes c0 = (1-beta)*c0 + beta*c, which moves a *static* body's
// stored centre by an ulp whenever SolveTOI advances it.  0 (default): statics are immovable (only
// alpha0 is updated) -- the documented deviation the CUDA path shares; 1: faithful ulp drift, used
// by tests to bound the effect.
extern int g_static_drift;

static inline void sincos_poly(double x, double* s, double* c) {
  const double kd = std::rint(x * 0.63661977236758134308);
  const long long k = (long long)kd;
  double r = (x - kd * 1.57079632673412561417e+00) - kd * 6.07710050650619224932e-11;
  const double z = r * r;
  const double ps =
      -1.66666666666666324348e-01 +
      z * (8.33333333332248946124e-03 +
           z * (-1.98412698298579493134e-04 +
                z * (2.75573137070700676789e-06 +
                     z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
  const double pc =
      4.16666666666666019037e-02 +
      z * (-1.38888888888741095749e-03 +
           z * (2.48015872894767294178e-05 +
                z * (-2.75573143513906633035e-07 +
                     z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
  const double sr = r + (r * z) * ps;
  const double cr = 1.0 - (0.5 * z - (z * z) * pc);
  switch ((int)(k & 3)) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}

static inline void sincosf_b2(float a, float* s, float* c) {
  if (g_trig_mode == 1) {
    *s = sinf(a);
    *c = cosf(a);
  } else {
    double sd, cd;
    sincos_poly((double)a, &sd, &cd);
    *s = (float)sd;
    *c = (float)cd;
  }
}

// ---- b2Math.h --------------------------------------------------------------------------------
struct V2 {
  float x, y;
};
static inline V2 mk(float x, float y) {
  V2 r;
  r.x = x;
  r.y = y;
  return r;
}
static inline V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
static inline V2 operator-(V2 a) { return mk(-a.x, -a.y); }
static inline V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
static inline void operator+=(V2& a, V2 b) {
  a.x += b.x;
  a.y += b.y;
}
static inline void operator-=(V2& a, V2 b) {
  a.x -= b.x;
  a.y -= b.y;
}
static inline void operator*=(V2& a, float s) {
  a.x *= s;
  a.y *= s;
}
static inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline V2 cross(V2 a, float s) { return mk(s * a.y, -s * a.x); }
static inline V2 cross(float s, V2 a) { return mk(-s * a.y, s * a.x); }
static inline float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline float lengthSq(V2 a) { return a.x * a.x + a.y * a.y; }
static inline float normalize(V2& a) {
  float len = length(a);
  if (len < kEps) return 0.0f;
  float inv = 1.0f / len;
  a.x *= inv;
  a.y *= inv;
  return len;
}
static inline float distance(V2 a, V2 b) { return length(a - b); }
static inline float distanceSq(V2 a, V2 b) {
  V2 c = a - b;
  return dot(c, c);
}
static inline float fmin2(float a, float b) { return a < b ? a : b; }
static inline float fmax2(float a, float b) { return a > b ? a : b; }
static inline float fclamp(float a, float lo, float hi) { return fmax2(lo, fmin2(a, hi)); }
static inline float fabs2(float a) { return a > 0.0f ? a : -a; }
static inline V2 vmin(V2 a, V2 b) { return mk(fmin2(a.x, b.x), fmin2(a.y, b.y)); }
static inline V2 vmax(V2 a, V2 b) { return mk(fmax2(a.x, b.x), fmax2(a.y, b.y)); }

struct Rot {
  float s, c;
  void set(float angle) { sincosf_b2(angle, &s, &c); }
};
struct Xf {
  V2 p;
  Rot q;
};
static inline V2 mul(Rot q, V2 v) { return mk(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
static inline V2 mulT(Rot q, V2 v) { return mk(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
static inline V2 mul(const Xf& T, V2 v) {
  float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
  float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
  return mk(x, y);
}
static inline V2 mulT(const Xf& T, V2 v) {
  float px = v.x - T.p.x;
  float py = v.y - T.p.y;
  float x = (T.q.c * px + T.q.s * py);
  float y = (-T.q.s * px + T.q.c * py);
  return mk(x, y);
}
struct Mat22 {
  V2 ex, ey;
  Mat22 inverse() const {
    float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
    Mat22 B;
    float det = a * d - b * c;
    if (det != 0.0f) det = 1.0f / det;
    B.ex.x = det * d;
    B.ey.x = -det * b;
    B.ex.y = -det * c;
    B.ey.y = det * a;
    return B;
  }
};
static inline V2 mul(const Mat22& A, V2 v) {
  return mk(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y);
}

struct Sweep {
  V2 localCenter, c0, c;
  float a0, a, alpha0;
  void getTransform(Xf* xf, float beta) const {
    xf->p = (1.0f - beta) * c0 + beta * c;
    float angle = (1.0f - beta) * a0 + beta * a;
    xf->q.set(angle);
    xf->p -= mul(xf->q, localCenter);
  }
  void advance(float alpha) {
    float beta = (alpha - alpha0) / (1.0f - alpha0);
    c0 = (1.0f - beta) * c0 + beta * c;
    a0 = (1.0f - beta) * a0 + beta * a;
    alpha0 = alpha;
  }
  void normalizeAngles() {
    float twoPi = 2.0f * kPi;
    float d = twoPi * floorf(a0 / twoPi);
    a0 -= d;
    a -= d;
  }
};

struct AABB {
  V2 lo, hi;
  bool contains(const AABB& o) const {
    bool r = true;
    r = r && lo.x <= o.lo.x;
    r = r && lo.y <= o.lo.y;
    r = r && o.hi.x <= hi.x;
    r = r && o.hi.y <= hi.y;
    return r;
  }
  void combine(const AABB& a, const AABB& b) {
    lo = vmin(a.lo, b.lo);
    hi = vmax(a.hi, b.hi);
  }
};
static inline bool testOverlap(const AABB& a, const AABB& b) {
  V2 d1 = b.lo - a.hi, d2 = a.lo - b.hi;
  if (d1.x > 0.0f || d1.y > 0.0f) return false;
  if (d2.x > 0.0f || d2.y > 0.0f) return false;
  return true;
}

// ---- shapes ----------------------------------------------------------------------------------
enum { SHAPE_CIRCLE = 0, SHAPE_POLYGON = 1 };
struct MassData {
  float mass;
  V2 center;
  float I;
};
struct Shape {
  int type;
  float radius;
  V2 p;  // circle centre (local)
  int count;
  V2 v[kMaxPolygonVertices], n[kMaxPolygonVertices];
  V2 centroid;

  void setCircle(float r) {
    type = SHAPE_CIRCLE;
    radius = r;
    p = mk(0, 0);
    count = 1;
    v[0] = p;
  }
  static V2 computeCentroid(const V2* vs, int count) {
    V2 c = mk(0.0f, 0.0f);
    float area = 0.0f;
    V2 pRef = mk(0.0f, 0.0f);
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < count; ++i) {
      V2 p1 = pRef, p2 = vs[i], p3 = i + 1 < count ? vs[i + 1] : vs[0];
      V2 e1 = p2 - p1, e2 = p3 - p1;
      float D = cross(e1, e2);
      float triangleArea = 0.5f * D;
      area += triangleArea;
      c += (triangleArea * inv3) * (p1 + p2 + p3);
    }
    c *= 1.0f / area;
    return c;
  }
  // b2PolygonShape::Set (2.3.0): weld, gift-wrap hull from the right-most point, normals, centroid.
  void setPolygon(const V2* vertices, int cnt) {
    type = SHAPE_POLYGON;
    radius = kPolygonRadius;
    p = mk(0, 0);
    int n_ = cnt < kMaxPolygonVertices ? cnt : kMaxPolygonVertices;
    V2 ps[kMaxPolygonVertices];
    int tempCount = 0;
    for (int i = 0; i < n_; ++i) {
      V2 vv = vertices[i];
      bool unique = true;
      for (int j = 0; j < tempCount; ++j)
        if (distanceSq(vv, ps[j]) < 0.5f * kLinearSlop) {
          unique = false;
          break;
        }
      if (unique) ps[tempCount++] = vv;
    }
    n_ = tempCount;
    int i0 = 0;
    float x0 = ps[0].x;
    for (int i = 1; i < n_; ++i) {
      float x = ps[i].x;
      if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) {
        i0 = i;
        x0 = x;
      }
    }
    int hull[kMaxPolygonVertices];
    int m = 0;
    int ih = i0;
    for (;;) {
      hull[m] = ih;
      int ie = 0;
      for (int j = 1; j < n_; ++j) {
        if (ie == ih) {
          ie = j;
          continue;
        }
        V2 r = ps[ie] - ps[hull[m]];
        V2 vv = ps[j] - ps[hull[m]];
        float c = cross(r, vv);
        if (c < 0.0f) ie = j;
        if (c == 0.0f && lengthSq(vv) > lengthSq(r)) ie = j;
      }
      ++m;
      ih = ie;
      if (ie == i0) break;
    }
    count = m;
    for (int i = 0; i < m; ++i) v[i] = ps[hull[i]];
    for (int i = 0; i < m; ++i) {
      int i1 = i, i2 = i + 1 < m ? i + 1 : 0;
      V2 edge = v[i2] - v[i1];
      n[i] = cross(edge, 1.0f);
      normalize(n[i]);
    }
    centroid = computeCentroid(v, m);
  }
  void computeAABB(AABB* aabb, const Xf& xf) const {
    if (type == SHAPE_CIRCLE) {
      V2 pp = xf.p + mul(xf.q, p);
      aabb->lo = mk(pp.x - radius, pp.y - radius);
      aabb->hi = mk(pp.x + radius, pp.y + radius);
      return;
    }
    V2 lower = mul(xf, v[0]);
    V2 upper = lower;
    for (int i = 1; i < count; ++i) {
      V2 vv = mul(xf, v[i]);
      lower = vmin(lower, vv);
      upper = vmax(upper, vv);
    }
    V2 r = mk(radius, radius);
    aabb->lo = lower - r;
    aabb->hi = upper + r;
  }
  void computeMass(MassData* md, float density) const {
    if (type == SHAPE_CIRCLE) {
      md->mass = density * kPi * radius * radius;
      md->center = p;
      md->I = md->mass * (0.5f * radius * radius + dot(p, p));
      return;
    }
    V2 center = mk(0.0f, 0.0f);
    float area = 0.0f, I = 0.0f;
    V2 s = mk(0.0f, 0.0f);
    for (int i = 0; i < count; ++i) s += v[i];
    s *= 1.0f / count;
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < count; ++i) {
      V2 e1 = v[i] - s;
      V2 e2 = i + 1 < count ? v[i + 1] - s : v[0] - s;
      float D = cross(e1, e2);
      float triangleArea = 0.5f * D;
      area += triangleArea;
      center += (triangleArea * k_inv3) * (e1 + e2);
      float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
      float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
      float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
      I += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    md->mass = density * area;
    center *= 1.0f / area;
    md->center = center + s;
    md->I = density * I;
    md->I += md->mass * (dot(md->center, md->center) - dot(center, center));
  }
};

// ---- manifolds (b2Collision.h) ---------------------------------------------------------------
enum { MANIFOLD_CIRCLES = 0, MANIFOLD_FACE_A = 1, MANIFOLD_FACE_B = 2 };
enum { FEATURE_VERTEX = 0, FEATURE_FACE = 1 };
static inline uint32_t makeKey(int indexA, int indexB, int typeA, int typeB) {
  return (uint32_t)(indexA & 255) | ((uint32_t)(indexB & 255) << 8) | ((uint32_t)(typeA & 255) << 16) |
         ((uint32_t)(typeB & 255) << 24);
}
struct ManifoldPoint {
  V2 localPoint;
  float normalImpulse, tangentImpulse;
  uint32_t key;
};
struct Manifold {
  ManifoldPoint points[2];
  V2 localNormal, localPoint;
  int type, pointCount;
};
struct ClipVertex {
  V2 v;
  uint32_t key;  // indexA | indexB<<8 | typeA<<16 | typeB<<24
};

void collidePolygonAndCircle(Manifold* m, const Shape* polyA, const Xf& xfA, const Shape* circB, const Xf& xfB);
void collidePolygons(Manifold* m, const Shape* polyA, const Xf& xfA, const Shape* polyB, const Xf& xfB);

// ---- GJK distance / TOI (b2Distance.cpp, b2TimeOfImpact.cpp) ----------------------------------
struct DistanceProxy {
  const V2* vertices;
  int count;
  float radius;
  void set(const Shape* s) {
    vertices = s->v;
    count = s->count;
    radius = s->radius;
    if (s->type == SHAPE_CIRCLE) {
      vertices = &s->p;
      count = 1;
    }
  }
  int getSupport(V2 d) const {
    int bestIndex = 0;
    float bestValue = dot(vertices[0], d);
    for (int i = 1; i < count; ++i) {
      float value = dot(vertices[i], d);
      if (value > bestValue) {
        bestIndex = i;
        bestValue = value;
      }
    }
    return bestIndex;
  }
};
struct SimplexCache {
  float metric;
  int count;
  int indexA[3], indexB[3];
};
struct DistanceInput {
  DistanceProxy proxyA, proxyB;
  Xf transformA, transformB;
  bool useRadii;
};
struct DistanceOutput {
  V2 pointA, pointB;
  float distance;
  int iterations;
};
void distanceGJK(DistanceOutput* out, SimplexCache* cache, const DistanceInput* in);
bool testOverlapShapes(const Shape* a, const Shape* b, const Xf& xfA, const Xf& xfB);

enum { TOI_UNKNOWN = 0, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };
struct TOIInput {
  DistanceProxy proxyA, proxyB;
  Sweep sweepA, sweepB;
  float tMax;
};
struct TOIOutput {
  int state;
  float t;
};
void timeOfImpact(TOIOutput* out, const TOIInput* in);

// ---- bodies, fixtures, contacts, world -------------------------------------------------------
enum { BODY_STATIC = 0, BODY_DYNAMIC = 2 };
struct Body {
  int type;
  Xf xf;
  Sweep sweep;
  V2 v;
  float w;
  V2 force;
  float torque;
  float mass, invMass, I, invI;
  float linearDamping, angularDamping;
  bool awake, islandFlag;
  float sleepTime;
  int islandIndex;
  int fixtureBegin, fixtureEnd;  // fixtures of this body are contiguous

  void synchronizeTransform() {
    xf.q.set(sweep.a);
    xf.p = sweep.c - mul(xf.q, sweep.localCenter);
  }
  void advance(float alpha) {
    if (type == BODY_STATIC && !g_static_drift) {
      sweep.alpha0 = alpha;
      return;
    }
    sweep.advance(alpha);
    sweep.c = sweep.c0;
    sweep.a = sweep.a0;
    xf.q.set(sweep.a);
    xf.p = sweep.c - mul(xf.q, sweep.localCenter);
  }
  void setAwake(bool flag) {
    if (flag) {
      if (!awake) {
        awake = true;
        sleepTime = 0.0f;
      }
    } else {
      awake = false;
      sleepTime = 0.0f;
      v = mk(0, 0);
      w = 0.0f;
      force = mk(0, 0);
      torque = 0.0f;
    }
  }
  void applyForceToCenter(V2 f, bool wake) {
    if (type != BODY_DYNAMIC) return;
    if (wake && !awake) setAwake(true);
    if (awake) force += f;
  }
  void applyTorque(float t, bool wake) {
    if (type != BODY_DYNAMIC) return;
    if (wake && !awake) setAwake(true);
    if (awake) torque += t;
  }
  void setLinearVelocity(V2 nv) {
    if (type == BODY_STATIC) return;
    if (dot(nv, nv) > 0.0f) setAwake(true);
    v = nv;
  }
};
struct Fixture {
  int body;
  Shape shape;
  float density, friction, restitution;
  unsigned categoryBits, maskBits;
  bool isSensor;
  AABB fatAABB;  // the dynamic-tree leaf box (b2DynamicTree::CreateProxy / MoveProxy)
};
struct Contact {
  int fA, fB;  // fixture indices; A/B as b2Contact::Create would order them
  Manifold manifold;
  bool touching, enabled, islandFlag, toiFlag;
  int toiCount;
  float toi;
  float friction, restitution;
};

struct World {
  std::vector<Body> bodies;        // creation order; Box2D's body list is the reverse of this
  std::vector<Fixture> fixtures;   // creation order == broad-phase proxy order in a fresh world
  std::vector<Contact*> contacts;  // world contact list, head first (newest first)
  std::vector<int> moveBuffer;     // b2BroadPhase move buffer (fixture indices)
  bool newFixture;                 // b2World::e_newFixture
  // listener: called with the contact on BeginContact (reference hockey_env.py:50-73)
  void (*beginContact)(void* user, const Contact* c);
  void* listenerUser;
  // statistics for tests
  long long nToiEvents, nToiCalls;

  World() : newFixture(false), beginContact(nullptr), listenerUser(nullptr), nToiEvents(0), nToiCalls(0) {}
  ~World() { clear(); }
  void clear();
  int createBody(int type, V2 position, float angle);
  int createFixture(int body, const Shape& shape, float density, float friction, float restitution,
                    unsigned cat, unsigned mask, bool sensor);
  void setTransform(int body, V2 position, float angle);
  void step(float dt, int velocityIterations, int positionIterations);
  Contact* findContact(int fA, int fB);
  void findNewContacts();

 private:
  void destroyContact(size_t idx);
  void updateContact(Contact* c);
  void collide();
  void solve(float dt, int velIters, int posIters);
  void solveTOI(float dt, int velIters);
  void synchronizeFixtures(int body);
  void moveProxy(int fixture, const AABB& aabb, V2 displacement);
  void clearForces();
};

}  // namespace b2mini
