// hk_collide.cuh -- narrow phase for the fixed HockeyEnv contact set: polygon-circle and
// polygon-polygon manifolds, GJK distance (sensor overlap) and conservative-advancement time of
// impact.  Replaces what the reference reaches through `self.world.Step(...)` (hockey_env.py:682)
// in Box2D's b2CollideCircle / b2CollidePolygon / b2Distance / b2TimeOfImpact; see DESIGN.md for the
// algorithm notes.  One env per thread: everything here is straight-line scalar code on registers
// plus small local arrays; polygon tables come from the Scene in constant/shared memory.
#pragma once
#include "hk_scene.cuh"

namespace hk {

enum { MANIFOLD_FACE_A = 1, MANIFOLD_FACE_B = 2 };
enum { FEATURE_VERTEX = 0, FEATURE_FACE = 1 };

HK_HD uint32_t makeKey(int indexA, int indexB, int typeA, int typeB) {
  return (uint32_t)(indexA & 255) | ((uint32_t)(indexB & 255) << 8) | ((uint32_t)(typeA & 255) << 16) |
         ((uint32_t)(typeB & 255) << 24);
}

struct Manifold {
  int type, count;
  float sepBound;  // lower bound of the core-shape distance at the evaluated poses: separation along sepNormal
  V2 sepNormal;    // face normal of polygon A (in A's frame) that achieves sepBound
  V2 localNormal, localPoint;
  V2 lp[2];
  uint32_t key[2];
  float ni[2], ti[2];
};

HK_HD V2 polyV(const Poly& p, int i) { return mk(p.vx[i], p.vy[i]); }
HK_HD V2 polyN(const Poly& p, int i) { return mk(p.nx[i], p.ny[i]); }

// ---- polygon (A) vs circle (B, centre at body origin) -------------------------------------------
HK_NI_NARROW void collidePolygonCircle(Manifold* m, const Poly& polyA, const Xf& xfA, V2 circleCenterWorld, float circleRadius) {
  m->count = 0;
  m->sepBound = -HK_MAXFLOAT;
  m->sepNormal = mk(0.0f, 0.0f);
  V2 cLocal = mulT(xfA, circleCenterWorld);
  int normalIndex = 0;
  float separation = -HK_MAXFLOAT;
  float radius = HK_POLYGON_RADIUS + circleRadius;
  int vertexCount = polyA.count;
  for (int i = 0; i < vertexCount; ++i) {
    float s = dot(polyN(polyA, i), cLocal - polyV(polyA, i));
    if (s > radius) {
      m->sepBound = s;
      m->sepNormal = polyN(polyA, i);
      return;
    }
    if (s > separation) {
      separation = s;
      normalIndex = i;
    }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
  V2 v1 = polyV(polyA, vertIndex1), v2 = polyV(polyA, vertIndex2);
  m->sepBound = separation;
  m->sepNormal = polyN(polyA, normalIndex);
  m->type = MANIFOLD_FACE_A;
  m->lp[0] = mk(0.0f, 0.0f);
  m->key[0] = 0;
  if (separation < HK_EPS) {
    m->count = 1;
    m->localNormal = polyN(polyA, normalIndex);
    m->localPoint = 0.5f * (v1 + v2);
    return;
  }
  float u1 = dot(cLocal - v1, v2 - v1);
  float u2 = dot(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (distanceSq(cLocal, v1) > radius * radius) return;
    m->count = 1;
    m->localNormal = cLocal - v1;
    normalize(m->localNormal);
    m->localPoint = v1;
  } else if (u2 <= 0.0f) {
    if (distanceSq(cLocal, v2) > radius * radius) return;
    m->count = 1;
    m->localNormal = cLocal - v2;
    normalize(m->localNormal);
    m->localPoint = v2;
  } else {
    V2 faceCenter = 0.5f * (v1 + v2);
    float sep = dot(cLocal - faceCenter, polyN(polyA, vertIndex1));
    if (sep > radius) return;
    m->count = 1;
    m->localNormal = polyN(polyA, vertIndex1);
    m->localPoint = faceCenter;
  }
}

// ---- polygon vs polygon (Box2D 2.3.0: hill-climbing max-separation search, 0.98/0.001 tie rule) --
HK_HD float edgeSeparation(const Poly& poly1, const Xf& xf1, int edge1, const Poly& poly2, const Xf& xf2) {
  V2 normal1World = mul(xf1.q, polyN(poly1, edge1));
  V2 normal1 = mulT(xf2.q, normal1World);
  int index = 0;
  float minDot = HK_MAXFLOAT;
  for (int i = 0; i < poly2.count; ++i) {
    float d = dot(polyV(poly2, i), normal1);
    if (d < minDot) {
      minDot = d;
      index = i;
    }
  }
  V2 v1 = mul(xf1, polyV(poly1, edge1));
  V2 v2 = mul(xf2, polyV(poly2, index));
  return dot(v2 - v1, normal1World);
}

HK_HD float findMaxSeparation(int* edgeIndex, const Poly& poly1, const Xf& xf1, const Poly& poly2, const Xf& xf2) {
  int count1 = poly1.count;
  V2 d = mul(xf2, mk(poly2.cenx, poly2.ceny)) - mul(xf1, mk(poly1.cenx, poly1.ceny));
  V2 dLocal1 = mulT(xf1.q, d);
  int edge = 0;
  float maxDot = -HK_MAXFLOAT;
  for (int i = 0; i < count1; ++i) {
    float dt = dot(polyN(poly1, i), dLocal1);
    if (dt > maxDot) {
      maxDot = dt;
      edge = i;
    }
  }
  float s = edgeSeparation(poly1, xf1, edge, poly2, xf2);
  int prevEdge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
  float sPrev = edgeSeparation(poly1, xf1, prevEdge, poly2, xf2);
  int nextEdge = edge + 1 < count1 ? edge + 1 : 0;
  float sNext = edgeSeparation(poly1, xf1, nextEdge, poly2, xf2);
  int bestEdge, increment;
  float bestSeparation;
  if (sPrev > s && sPrev > sNext) {
    increment = -1;
    bestEdge = prevEdge;
    bestSeparation = sPrev;
  } else if (sNext > s) {
    increment = 1;
    bestEdge = nextEdge;
    bestSeparation = sNext;
  } else {
    *edgeIndex = edge;
    return s;
  }
  for (;;) {
    if (increment == -1)
      edge = bestEdge - 1 >= 0 ? bestEdge - 1 : count1 - 1;
    else
      edge = bestEdge + 1 < count1 ? bestEdge + 1 : 0;
    s = edgeSeparation(poly1, xf1, edge, poly2, xf2);
    if (s > bestSeparation) {
      bestEdge = edge;
      bestSeparation = s;
    } else {
      break;
    }
  }
  *edgeIndex = bestEdge;
  return bestSeparation;
}

struct ClipVertex {
  V2 v;
  uint32_t key;
};

HK_HD int clipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], V2 normal, float offset, int vertexIndexA) {
  int numOut = 0;
  float distance0 = dot(normal, vIn[0].v) - offset;
  float distance1 = dot(normal, vIn[1].v) - offset;
  if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
  if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
  if (distance0 * distance1 < 0.0f) {
    float interp = distance0 / (distance0 - distance1);
    vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
    int indexB = (int)((vIn[0].key >> 8) & 255);
    vOut[numOut].key = makeKey(vertexIndexA, indexB, FEATURE_VERTEX, FEATURE_FACE);
    ++numOut;
  }
  return numOut;
}

HK_NI_NARROW void collidePolygons(Manifold* m, const Poly& polyA, const Xf& xfA, const Poly& polyB, const Xf& xfB) {
  m->count = 0;
  const float totalRadius = HK_POLYGON_RADIUS + HK_POLYGON_RADIUS;
  int edgeA = 0;
  float separationA = findMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
  // a face separation of A is a lower bound of the distance between the convex cores, along that face normal
  m->sepBound = separationA;
  m->sepNormal = polyN(polyA, edgeA);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = findMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
  if (separationB > m->sepBound) {  // a face separation of B bounds the distance as well, but its normal moves with B
    m->sepBound = separationB;
    m->sepNormal = mk(0.0f, 0.0f);
  }
  if (separationB > totalRadius) return;
  const float k_relativeTol = 0.98f;
  const float k_absoluteTol = 0.001f;
  const bool flip = separationB > k_relativeTol * separationA + k_absoluteTol;
  const Poly& poly1 = flip ? polyB : polyA;
  const Poly& poly2 = flip ? polyA : polyB;
  const Xf xf1 = flip ? xfB : xfA;
  const Xf xf2 = flip ? xfA : xfB;
  const int edge1 = flip ? edgeB : edgeA;
  m->type = flip ? MANIFOLD_FACE_B : MANIFOLD_FACE_A;

  // b2FindIncidentEdge
  ClipVertex incidentEdge[2];
  {
    V2 normal1 = mulT(xf2.q, mul(xf1.q, polyN(poly1, edge1)));
    int index = 0;
    float minDot = HK_MAXFLOAT;
    for (int i = 0; i < poly2.count; ++i) {
      float d = dot(normal1, polyN(poly2, i));
      if (d < minDot) {
        minDot = d;
        index = i;
      }
    }
    int i1 = index;
    int i2 = i1 + 1 < poly2.count ? i1 + 1 : 0;
    incidentEdge[0].v = mul(xf2, polyV(poly2, i1));
    incidentEdge[0].key = makeKey(edge1, i1, FEATURE_FACE, FEATURE_VERTEX);
    incidentEdge[1].v = mul(xf2, polyV(poly2, i2));
    incidentEdge[1].key = makeKey(edge1, i2, FEATURE_FACE, FEATURE_VERTEX);
  }
  int count1 = poly1.count;
  int iv1 = edge1;
  int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  V2 v11 = polyV(poly1, iv1), v12 = polyV(poly1, iv2);
  V2 localTangent = v12 - v11;
  normalize(localTangent);
  V2 localNormal = cross(localTangent, 1.0f);
  V2 planePoint = 0.5f * (v11 + v12);
  V2 tangent = mul(xf1.q, localTangent);
  V2 normal = cross(tangent, 1.0f);
  v11 = mul(xf1, v11);
  v12 = mul(xf1, v12);
  float frontOffset = dot(normal, v11);
  float sideOffset1 = -dot(tangent, v11) + totalRadius;
  float sideOffset2 = dot(tangent, v12) + totalRadius;
  ClipVertex clipPoints1[2], clipPoints2[2];
  int np = clipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = clipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  m->localNormal = localNormal;
  m->localPoint = planePoint;
  int pointCount = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = dot(normal, clipPoints2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      m->lp[pointCount] = mulT(xf2, clipPoints2[i].v);
      uint32_t k = clipPoints2[i].key;
      if (flip) {
        int iA = k & 255, iB = (k >> 8) & 255, tA = (k >> 16) & 255, tB = (k >> 24) & 255;
        k = makeKey(iB, iA, tB, tA);
      }
      m->key[pointCount] = k;
      ++pointCount;
    }
  }
  m->count = pointCount;
}

// ---- GJK (b2Distance) ---------------------------------------------------------------------------
// A proxy is either a scene polygon or the puck centre (one vertex at the local origin).
struct Proxy {
  const Poly* poly;  // nullptr => single point (0,0)
  float radius;
};
HK_HD int proxyCount(const Proxy& p) { return p.poly ? p.poly->count : 1; }
HK_HD V2 proxyVertex(const Proxy& p, int i) { return p.poly ? polyV(*p.poly, i) : mk(0.0f, 0.0f); }
HK_HD int proxySupport(const Proxy& p, V2 d) {
  if (!p.poly) return 0;
  int bestIndex = 0;
  float bestValue = dot(polyV(*p.poly, 0), d);
  for (int i = 1; i < p.poly->count; ++i) {
    float value = dot(polyV(*p.poly, i), d);
    if (value > bestValue) {
      bestIndex = i;
      bestValue = value;
    }
  }
  return bestIndex;
}

struct SimplexCache {
  float metric;
  int count;
  int indexA[3], indexB[3];
};
struct SimplexVertex {
  V2 wA, wB, w;
  float a;
  int indexA, indexB;
};
struct Simplex {
  SimplexVertex v[3];
  int count;
};

HK_HD float simplexMetric(const Simplex& s) {
  switch (s.count) {
    case 2:
      return length(s.v[0].w - s.v[1].w);
    case 3:
      return cross(s.v[1].w - s.v[0].w, s.v[2].w - s.v[0].w);
    default:
      return 0.0f;
  }
}

HK_HD void simplexSolve2(Simplex& s) {
  V2 w1 = s.v[0].w, w2 = s.v[1].w;
  V2 e12 = w2 - w1;
  float d12_2 = -dot(w1, e12);
  if (d12_2 <= 0.0f) {
    s.v[0].a = 1.0f;
    s.count = 1;
    return;
  }
  float d12_1 = dot(w2, e12);
  if (d12_1 <= 0.0f) {
    s.v[1].a = 1.0f;
    s.count = 1;
    s.v[0] = s.v[1];
    return;
  }
  float inv_d12 = 1.0f / (d12_1 + d12_2);
  s.v[0].a = d12_1 * inv_d12;
  s.v[1].a = d12_2 * inv_d12;
  s.count = 2;
}

HK_HD void simplexSolve3(Simplex& s) {
  V2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
  V2 e12 = w2 - w1;
  float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12);
  float d12_1 = w2e12, d12_2 = -w1e12;
  V2 e13 = w3 - w1;
  float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13);
  float d13_1 = w3e13, d13_2 = -w1e13;
  V2 e23 = w3 - w2;
  float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23);
  float d23_1 = w3e23, d23_2 = -w2e23;
  float n123 = cross(e12, e13);
  float d123_1 = n123 * cross(w2, w3);
  float d123_2 = n123 * cross(w3, w1);
  float d123_3 = n123 * cross(w1, w2);
  if (d12_2 <= 0.0f && d13_2 <= 0.0f) {
    s.v[0].a = 1.0f;
    s.count = 1;
    return;
  }
  if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
    float inv_d12 = 1.0f / (d12_1 + d12_2);
    s.v[0].a = d12_1 * inv_d12;
    s.v[1].a = d12_2 * inv_d12;
    s.count = 2;
    return;
  }
  if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
    float inv_d13 = 1.0f / (d13_1 + d13_2);
    s.v[0].a = d13_1 * inv_d13;
    s.v[2].a = d13_2 * inv_d13;
    s.count = 2;
    s.v[1] = s.v[2];
    return;
  }
  if (d12_1 <= 0.0f && d23_2 <= 0.0f) {
    s.v[1].a = 1.0f;
    s.count = 1;
    s.v[0] = s.v[1];
    return;
  }
  if (d13_1 <= 0.0f && d23_1 <= 0.0f) {
    s.v[2].a = 1.0f;
    s.count = 1;
    s.v[0] = s.v[2];
    return;
  }
  if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
    float inv_d23 = 1.0f / (d23_1 + d23_2);
    s.v[1].a = d23_1 * inv_d23;
    s.v[2].a = d23_2 * inv_d23;
    s.count = 2;
    s.v[0] = s.v[2];
    return;
  }
  float inv_d123 = 1.0f / (d123_1 + d123_2 + d123_3);
  s.v[0].a = d123_1 * inv_d123;
  s.v[1].a = d123_2 * inv_d123;
  s.v[2].a = d123_3 * inv_d123;
  s.count = 3;
}

// returns the core distance (useRadii handled by the caller); fills witness points
HK_NI_TOI float gjkDistance(SimplexCache* cache, const Proxy& proxyA, const Xf& xfA, const Proxy& proxyB,
                                 const Xf& xfB, V2* pointA, V2* pointB) {
  Simplex simplex;
  // ReadCache
  simplex.count = cache->count;
  for (int i = 0; i < simplex.count; ++i) {
    SimplexVertex* sv = simplex.v + i;
    sv->indexA = cache->indexA[i];
    sv->indexB = cache->indexB[i];
    sv->wA = mul(xfA, proxyVertex(proxyA, sv->indexA));
    sv->wB = mul(xfB, proxyVertex(proxyB, sv->indexB));
    sv->w = sv->wB - sv->wA;
    sv->a = 0.0f;
  }
  if (simplex.count > 1) {
    float metric1 = cache->metric;
    float metric2 = simplexMetric(simplex);
    if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < HK_EPS) simplex.count = 0;
  }
  if (simplex.count == 0) {
    SimplexVertex* sv = simplex.v + 0;
    sv->indexA = 0;
    sv->indexB = 0;
    sv->wA = mul(xfA, proxyVertex(proxyA, 0));
    sv->wB = mul(xfB, proxyVertex(proxyB, 0));
    sv->w = sv->wB - sv->wA;
    sv->a = 1.0f;
    simplex.count = 1;
  }
  const int k_maxIters = 20;
  int saveA[3], saveB[3];
  int iter = 0;
  while (iter < k_maxIters) {
    int saveCount = simplex.count;
    for (int i = 0; i < saveCount; ++i) {
      saveA[i] = simplex.v[i].indexA;
      saveB[i] = simplex.v[i].indexB;
    }
    if (simplex.count == 2) simplexSolve2(simplex);
    else if (simplex.count == 3) simplexSolve3(simplex);
    if (simplex.count == 3) break;
    V2 d;
    if (simplex.count == 1) {
      d = -simplex.v[0].w;
    } else {
      V2 e12 = simplex.v[1].w - simplex.v[0].w;
      float sgn = cross(e12, -simplex.v[0].w);
      d = sgn > 0.0f ? cross(1.0f, e12) : cross(e12, 1.0f);
    }
    if (lengthSq(d) < HK_EPS * HK_EPS) break;
    SimplexVertex* vertex = simplex.v + simplex.count;
    vertex->indexA = proxySupport(proxyA, mulT(xfA.q, -d));
    vertex->wA = mul(xfA, proxyVertex(proxyA, vertex->indexA));
    vertex->indexB = proxySupport(proxyB, mulT(xfB.q, d));
    vertex->wB = mul(xfB, proxyVertex(proxyB, vertex->indexB));
    vertex->w = vertex->wB - vertex->wA;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < saveCount; ++i)
      if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) duplicate = true;
    if (duplicate) break;
    ++simplex.count;
  }
  // witness points
  if (simplex.count == 1) {
    *pointA = simplex.v[0].wA;
    *pointB = simplex.v[0].wB;
  } else if (simplex.count == 2) {
    *pointA = simplex.v[0].a * simplex.v[0].wA + simplex.v[1].a * simplex.v[1].wA;
    *pointB = simplex.v[0].a * simplex.v[0].wB + simplex.v[1].a * simplex.v[1].wB;
  } else {
    *pointA = simplex.v[0].a * simplex.v[0].wA + simplex.v[1].a * simplex.v[1].wA + simplex.v[2].a * simplex.v[2].wA;
    *pointB = *pointA;
  }
  float dist = length(*pointA - *pointB);
  cache->metric = simplexMetric(simplex);
  cache->count = simplex.count;
  for (int i = 0; i < simplex.count; ++i) {
    cache->indexA[i] = simplex.v[i].indexA;
    cache->indexB[i] = simplex.v[i].indexB;
  }
  return dist;
}

// b2TestOverlap(shapeA, shapeB, xfA, xfB): sensor test of b2Contact::Update (goal polygon vs puck)
HK_NI_TOI bool testOverlapPolyPuck(const Poly& poly, const Xf& xfA, V2 puckCenter, float puckRadius) {
  Proxy pa, pb;
  pa.poly = &poly;
  pa.radius = HK_POLYGON_RADIUS;
  pb.poly = nullptr;
  pb.radius = puckRadius;
  Xf xfB;
  xfB.p = puckCenter;
  xfB.q.s = 0.0f;  // the rotation of a circle centred on its body origin never enters the result
  xfB.q.c = 1.0f;
  SimplexCache cache;
  cache.count = 0;
  V2 a, b;
  float d = gjkDistance(&cache, pa, xfA, pb, xfB, &a, &b);
  float rA = pa.radius, rB = pb.radius;
  if (d > rA + rB && d > HK_EPS) d -= rA + rB;
  else d = 0.0f;
  return d < 10.0f * HK_EPS;
}

// ---- b2TimeOfImpact ------------------------------------------------------------------------------
enum { SEP_POINTS = 0, SEP_FACE_A = 1, SEP_FACE_B = 2 };
enum { TOI_UNKNOWN = 0, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };

struct SepFn {
  int type;
  V2 localPoint, axis;
};

HK_NI_TOI float sepFindMin(const SepFn& f, const Proxy& pA, const Sweep& sA, const Proxy& pB, const Sweep& sB, int* indexA,
                       int* indexB, float t) {
  Xf xfA, xfB;
  sweepXf(sA, &xfA, t);
  sweepXf(sB, &xfB, t);
  if (f.type == SEP_POINTS) {
    V2 axisA = mulT(xfA.q, f.axis);
    V2 axisB = mulT(xfB.q, -f.axis);
    *indexA = proxySupport(pA, axisA);
    *indexB = proxySupport(pB, axisB);
    V2 pointA = mul(xfA, proxyVertex(pA, *indexA));
    V2 pointB = mul(xfB, proxyVertex(pB, *indexB));
    return dot(pointB - pointA, f.axis);
  } else if (f.type == SEP_FACE_A) {
    V2 normal = mul(xfA.q, f.axis);
    V2 pointA = mul(xfA, f.localPoint);
    V2 axisB = mulT(xfB.q, -normal);
    *indexA = -1;
    *indexB = proxySupport(pB, axisB);
    V2 pointB = mul(xfB, proxyVertex(pB, *indexB));
    return dot(pointB - pointA, normal);
  } else {
    V2 normal = mul(xfB.q, f.axis);
    V2 pointB = mul(xfB, f.localPoint);
    V2 axisA = mulT(xfA.q, -normal);
    *indexB = -1;
    *indexA = proxySupport(pA, axisA);
    V2 pointA = mul(xfA, proxyVertex(pA, *indexA));
    return dot(pointA - pointB, normal);
  }
}

HK_NI_TOI float sepEvaluate(const SepFn& f, const Proxy& pA, const Sweep& sA, const Proxy& pB, const Sweep& sB, int indexA,
                        int indexB, float t) {
  Xf xfA, xfB;
  sweepXf(sA, &xfA, t);
  sweepXf(sB, &xfB, t);
  if (f.type == SEP_POINTS) {
    V2 pointA = mul(xfA, proxyVertex(pA, indexA));
    V2 pointB = mul(xfB, proxyVertex(pB, indexB));
    return dot(pointB - pointA, f.axis);
  } else if (f.type == SEP_FACE_A) {
    V2 normal = mul(xfA.q, f.axis);
    V2 pointA = mul(xfA, f.localPoint);
    V2 pointB = mul(xfB, proxyVertex(pB, indexB));
    return dot(pointB - pointA, normal);
  } else {
    V2 normal = mul(xfB.q, f.axis);
    V2 pointB = mul(xfB, f.localPoint);
    V2 pointA = mul(xfA, proxyVertex(pA, indexA));
    return dot(pointA - pointB, normal);
  }
}

HK_NI_TOIFN void timeOfImpact(int* outState, float* outT, const Proxy& proxyA, const Sweep& sweepAin,
                                 const Proxy& proxyB, const Sweep& sweepBin, float tMax) {
  *outState = TOI_UNKNOWN;
  *outT = tMax;
  Sweep sweepA = sweepAin, sweepB = sweepBin;
  sweepNormalize(sweepA);
  sweepNormalize(sweepB);
  float totalRadius = proxyA.radius + proxyB.radius;
  float target = fmax2(HK_LINEAR_SLOP, totalRadius - 3.0f * HK_LINEAR_SLOP);
  float tolerance = 0.25f * HK_LINEAR_SLOP;
  float t1 = 0.0f;
  const int k_maxIterations = 20;
  int iter = 0;
  SimplexCache cache;
  cache.count = 0;
  for (;;) {
    Xf xfA, xfB;
    sweepXf(sweepA, &xfA, t1);
    sweepXf(sweepB, &xfB, t1);
    V2 wa, wb;
    float dist = gjkDistance(&cache, proxyA, xfA, proxyB, xfB, &wa, &wb);
    if (dist <= 0.0f) {
      *outState = TOI_OVERLAPPED;
      *outT = 0.0f;
      break;
    }
    if (dist < target + tolerance) {
      *outState = TOI_TOUCHING;
      *outT = t1;
      break;
    }
    // b2SeparationFunction::Initialize (xfA/xfB at t1 are the transforms just computed)
    SepFn fcn;
    if (cache.count == 1) {
      fcn.type = SEP_POINTS;
      V2 pointA = mul(xfA, proxyVertex(proxyA, cache.indexA[0]));
      V2 pointB = mul(xfB, proxyVertex(proxyB, cache.indexB[0]));
      fcn.axis = pointB - pointA;
      normalize(fcn.axis);
      fcn.localPoint = mk(0.0f, 0.0f);
    } else if (cache.indexA[0] == cache.indexA[1]) {
      fcn.type = SEP_FACE_B;
      V2 localPointB1 = proxyVertex(proxyB, cache.indexB[0]);
      V2 localPointB2 = proxyVertex(proxyB, cache.indexB[1]);
      fcn.axis = cross(localPointB2 - localPointB1, 1.0f);
      normalize(fcn.axis);
      V2 normal = mul(xfB.q, fcn.axis);
      fcn.localPoint = 0.5f * (localPointB1 + localPointB2);
      V2 pointB = mul(xfB, fcn.localPoint);
      V2 pointA = mul(xfA, proxyVertex(proxyA, cache.indexA[0]));
      float s = dot(pointA - pointB, normal);
      if (s < 0.0f) fcn.axis = -fcn.axis;
    } else {
      fcn.type = SEP_FACE_A;
      V2 localPointA1 = proxyVertex(proxyA, cache.indexA[0]);
      V2 localPointA2 = proxyVertex(proxyA, cache.indexA[1]);
      fcn.axis = cross(localPointA2 - localPointA1, 1.0f);
      normalize(fcn.axis);
      V2 normal = mul(xfA.q, fcn.axis);
      fcn.localPoint = 0.5f * (localPointA1 + localPointA2);
      V2 pointA = mul(xfA, fcn.localPoint);
      V2 pointB = mul(xfB, proxyVertex(proxyB, cache.indexB[0]));
      float s = dot(pointB - pointA, normal);
      if (s < 0.0f) fcn.axis = -fcn.axis;
    }
    bool done = false;
    float t2 = tMax;
    int pushBackIter = 0;
    for (;;) {
      int indexA, indexB;
      float s2 = sepFindMin(fcn, proxyA, sweepA, proxyB, sweepB, &indexA, &indexB, t2);
      if (s2 > target + tolerance) {
        *outState = TOI_SEPARATED;
        *outT = tMax;
        done = true;
        break;
      }
      if (s2 > target - tolerance) {
        t1 = t2;
        break;
      }
      float s1 = sepEvaluate(fcn, proxyA, sweepA, proxyB, sweepB, indexA, indexB, t1);
      if (s1 < target - tolerance) {
        *outState = TOI_FAILED;
        *outT = t1;
        done = true;
        break;
      }
      if (s1 <= target + tolerance) {
        *outState = TOI_TOUCHING;
        *outT = t1;
        done = true;
        break;
      }
      int rootIterCount = 0;
      float a1 = t1, a2 = t2;
      for (;;) {
        float t;
        if (rootIterCount & 1)
          t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
        else
          t = 0.5f * (a1 + a2);
        ++rootIterCount;
        float s = sepEvaluate(fcn, proxyA, sweepA, proxyB, sweepB, indexA, indexB, t);
        if (fabs2(s - target) < tolerance) {
          t2 = t;
          break;
        }
        if (s > target) {
          a1 = t;
          s1 = s;
        } else {
          a2 = t;
          s2 = s;
        }
        if (rootIterCount == 50) break;
      }
      ++pushBackIter;
      if (pushBackIter == HK_MAX_POLY_VERTS) break;
    }
    ++iter;
    if (done) break;
    if (iter == k_maxIterations) {
      *outState = TOI_FAILED;
      *outT = t1;
      break;
    }
  }
}

}  // namespace hk
