// oracle/b2mini.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See b2mini.h for the header note.
// Restates (function by function) the Box2D 2.3.0 algorithms reached from the reference call
// `self.world.Step(self.timeStep, 6 * 30, 2 * 30)` (hockey/hockey_env.py:682).
#include "b2mini.h"

namespace b2mini {

int g_trig_mode = 0;
int g_static_drift = 0;

// =============================== b2CollideCircle.cpp ==========================================
void collidePolygonAndCircle(Manifold* manifold, const Shape* polygonA, const Xf& xfA, const Shape* circleB,
                             const Xf& xfB) {
  manifold->pointCount = 0;
  V2 c = mul(xfB, circleB->p);
  V2 cLocal = mulT(xfA, c);
  int normalIndex = 0;
  float separation = -kMaxFloat;
  float radius = polygonA->radius + circleB->radius;
  int vertexCount = polygonA->count;
  const V2* vertices = polygonA->v;
  const V2* normals = polygonA->n;
  for (int i = 0; i < vertexCount; ++i) {
    float s = dot(normals[i], cLocal - vertices[i]);
    if (s > radius) return;
    if (s > separation) {
      separation = s;
      normalIndex = i;
    }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
  V2 v1 = vertices[vertIndex1], v2 = vertices[vertIndex2];
  if (separation < kEps) {
    manifold->pointCount = 1;
    manifold->type = MANIFOLD_FACE_A;
    manifold->localNormal = normals[normalIndex];
    manifold->localPoint = 0.5f * (v1 + v2);
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].key = 0;
    return;
  }
  float u1 = dot(cLocal - v1, v2 - v1);
  float u2 = dot(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (distanceSq(cLocal, v1) > radius * radius) return;
    manifold->pointCount = 1;
    manifold->type = MANIFOLD_FACE_A;
    manifold->localNormal = cLocal - v1;
    normalize(manifold->localNormal);
    manifold->localPoint = v1;
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].key = 0;
  } else if (u2 <= 0.0f) {
    if (distanceSq(cLocal, v2) > radius * radius) return;
    manifold->pointCount = 1;
    manifold->type = MANIFOLD_FACE_A;
    manifold->localNormal = cLocal - v2;
    normalize(manifold->localNormal);
    manifold->localPoint = v2;
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].key = 0;
  } else {
    V2 faceCenter = 0.5f * (v1 + v2);
    float sep = dot(cLocal - faceCenter, normals[vertIndex1]);
    if (sep > radius) return;
    manifold->pointCount = 1;
    manifold->type = MANIFOLD_FACE_A;
    manifold->localNormal = normals[vertIndex1];
    manifold->localPoint = faceCenter;
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].key = 0;
  }
}

// =============================== b2CollidePolygon.cpp (2.3.0) =================================
static float edgeSeparation(const Shape* poly1, const Xf& xf1, int edge1, const Shape* poly2, const Xf& xf2) {
  const V2* vertices1 = poly1->v;
  const V2* normals1 = poly1->n;
  int count2 = poly2->count;
  const V2* vertices2 = poly2->v;
  V2 normal1World = mul(xf1.q, normals1[edge1]);
  V2 normal1 = mulT(xf2.q, normal1World);
  int index = 0;
  float minDot = kMaxFloat;
  for (int i = 0; i < count2; ++i) {
    float d = dot(vertices2[i], normal1);
    if (d < minDot) {
      minDot = d;
      index = i;
    }
  }
  V2 v1 = mul(xf1, vertices1[edge1]);
  V2 v2 = mul(xf2, vertices2[index]);
  return dot(v2 - v1, normal1World);
}

static float findMaxSeparation(int* edgeIndex, const Shape* poly1, const Xf& xf1, const Shape* poly2,
                               const Xf& xf2) {
  int count1 = poly1->count;
  const V2* normals1 = poly1->n;
  V2 d = mul(xf2, poly2->centroid) - mul(xf1, poly1->centroid);
  V2 dLocal1 = mulT(xf1.q, d);
  int edge = 0;
  float maxDot = -kMaxFloat;
  for (int i = 0; i < count1; ++i) {
    float dt = dot(normals1[i], dLocal1);
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
  int bestEdge;
  float bestSeparation;
  int increment;
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

static void findIncidentEdge(ClipVertex c[2], const Shape* poly1, const Xf& xf1, int edge1, const Shape* poly2,
                             const Xf& xf2) {
  const V2* normals1 = poly1->n;
  int count2 = poly2->count;
  const V2* vertices2 = poly2->v;
  const V2* normals2 = poly2->n;
  V2 normal1 = mulT(xf2.q, mul(xf1.q, normals1[edge1]));
  int index = 0;
  float minDot = kMaxFloat;
  for (int i = 0; i < count2; ++i) {
    float d = dot(normal1, normals2[i]);
    if (d < minDot) {
      minDot = d;
      index = i;
    }
  }
  int i1 = index;
  int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
  c[0].v = mul(xf2, vertices2[i1]);
  c[0].key = makeKey(edge1, i1, FEATURE_FACE, FEATURE_VERTEX);
  c[1].v = mul(xf2, vertices2[i2]);
  c[1].key = makeKey(edge1, i2, FEATURE_FACE, FEATURE_VERTEX);
}

static int clipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], V2 normal, float offset,
                             int vertexIndexA) {
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

void collidePolygons(Manifold* manifold, const Shape* polyA, const Xf& xfA, const Shape* polyB, const Xf& xfB) {
  manifold->pointCount = 0;
  float totalRadius = polyA->radius + polyB->radius;
  int edgeA = 0;
  float separationA = findMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = findMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
  if (separationB > totalRadius) return;

  const Shape* poly1;
  const Shape* poly2;
  Xf xf1, xf2;
  int edge1;
  int flip;
  const float k_relativeTol = 0.98f;
  const float k_absoluteTol = 0.001f;
  if (separationB > k_relativeTol * separationA + k_absoluteTol) {
    poly1 = polyB;
    poly2 = polyA;
    xf1 = xfB;
    xf2 = xfA;
    edge1 = edgeB;
    manifold->type = MANIFOLD_FACE_B;
    flip = 1;
  } else {
    poly1 = polyA;
    poly2 = polyB;
    xf1 = xfA;
    xf2 = xfB;
    edge1 = edgeA;
    manifold->type = MANIFOLD_FACE_A;
    flip = 0;
  }
  ClipVertex incidentEdge[2];
  findIncidentEdge(incidentEdge, poly1, xf1, edge1, poly2, xf2);
  int count1 = poly1->count;
  const V2* vertices1 = poly1->v;
  int iv1 = edge1;
  int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  V2 v11 = vertices1[iv1], v12 = vertices1[iv2];
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
  int np;
  np = clipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = clipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  manifold->localNormal = localNormal;
  manifold->localPoint = planePoint;
  int pointCount = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = dot(normal, clipPoints2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      ManifoldPoint* cp = manifold->points + pointCount;
      cp->localPoint = mulT(xf2, clipPoints2[i].v);
      cp->key = clipPoints2[i].key;
      if (flip) {
        uint32_t k = cp->key;
        int iA = k & 255, iB = (k >> 8) & 255, tA = (k >> 16) & 255, tB = (k >> 24) & 255;
        cp->key = makeKey(iB, iA, tB, tA);
      }
      ++pointCount;
    }
  }
  manifold->pointCount = pointCount;
}

// =============================== b2Distance.cpp ================================================
struct SimplexVertex {
  V2 wA, wB, w;
  float a;
  int indexA, indexB;
};
struct Simplex {
  SimplexVertex v[3];
  int count;

  void readCache(const SimplexCache* cache, const DistanceProxy* proxyA, const Xf& xfA, const DistanceProxy* proxyB,
                 const Xf& xfB) {
    count = cache->count;
    for (int i = 0; i < count; ++i) {
      SimplexVertex* sv = v + i;
      sv->indexA = cache->indexA[i];
      sv->indexB = cache->indexB[i];
      V2 wALocal = proxyA->vertices[sv->indexA];
      V2 wBLocal = proxyB->vertices[sv->indexB];
      sv->wA = mul(xfA, wALocal);
      sv->wB = mul(xfB, wBLocal);
      sv->w = sv->wB - sv->wA;
      sv->a = 0.0f;
    }
    if (count > 1) {
      float metric1 = cache->metric;
      float metric2 = getMetric();
      if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEps) count = 0;
    }
    if (count == 0) {
      SimplexVertex* sv = v + 0;
      sv->indexA = 0;
      sv->indexB = 0;
      V2 wALocal = proxyA->vertices[0];
      V2 wBLocal = proxyB->vertices[0];
      sv->wA = mul(xfA, wALocal);
      sv->wB = mul(xfB, wBLocal);
      sv->w = sv->wB - sv->wA;
      sv->a = 1.0f;
      count = 1;
    }
  }
  void writeCache(SimplexCache* cache) const {
    cache->metric = getMetric();
    cache->count = count;
    for (int i = 0; i < count; ++i) {
      cache->indexA[i] = v[i].indexA;
      cache->indexB[i] = v[i].indexB;
    }
  }
  V2 getSearchDirection() const {
    switch (count) {
      case 1:
        return -v[0].w;
      case 2: {
        V2 e12 = v[1].w - v[0].w;
        float sgn = cross(e12, -v[0].w);
        if (sgn > 0.0f) return cross(1.0f, e12);
        return cross(e12, 1.0f);
      }
      default:
        return mk(0, 0);
    }
  }
  void getWitnessPoints(V2* pA, V2* pB) const {
    switch (count) {
      case 1:
        *pA = v[0].wA;
        *pB = v[0].wB;
        break;
      case 2:
        *pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
        *pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
        break;
      case 3:
        *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
        *pB = *pA;
        break;
      default:
        break;
    }
  }
  float getMetric() const {
    switch (count) {
      case 1:
        return 0.0f;
      case 2:
        return distance(v[0].w, v[1].w);
      case 3:
        return cross(v[1].w - v[0].w, v[2].w - v[0].w);
      default:
        return 0.0f;
    }
  }
  void solve2() {
    V2 w1 = v[0].w, w2 = v[1].w;
    V2 e12 = w2 - w1;
    float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) {
      v[0].a = 1.0f;
      count = 1;
      return;
    }
    float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) {
      v[1].a = 1.0f;
      count = 1;
      v[0] = v[1];
      return;
    }
    float inv_d12 = 1.0f / (d12_1 + d12_2);
    v[0].a = d12_1 * inv_d12;
    v[1].a = d12_2 * inv_d12;
    count = 2;
  }
  void solve3() {
    V2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
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
      v[0].a = 1.0f;
      count = 1;
      return;
    }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
      float inv_d12 = 1.0f / (d12_1 + d12_2);
      v[0].a = d12_1 * inv_d12;
      v[1].a = d12_2 * inv_d12;
      count = 2;
      return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
      float inv_d13 = 1.0f / (d13_1 + d13_2);
      v[0].a = d13_1 * inv_d13;
      v[2].a = d13_2 * inv_d13;
      count = 2;
      v[1] = v[2];
      return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) {
      v[1].a = 1.0f;
      count = 1;
      v[0] = v[1];
      return;
    }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) {
      v[2].a = 1.0f;
      count = 1;
      v[0] = v[2];
      return;
    }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
      float inv_d23 = 1.0f / (d23_1 + d23_2);
      v[1].a = d23_1 * inv_d23;
      v[2].a = d23_2 * inv_d23;
      count = 2;
      v[0] = v[2];
      return;
    }
    float inv_d123 = 1.0f / (d123_1 + d123_2 + d123_3);
    v[0].a = d123_1 * inv_d123;
    v[1].a = d123_2 * inv_d123;
    v[2].a = d123_3 * inv_d123;
    count = 3;
  }
};

void distanceGJK(DistanceOutput* output, SimplexCache* cache, const DistanceInput* input) {
  const DistanceProxy* proxyA = &input->proxyA;
  const DistanceProxy* proxyB = &input->proxyB;
  Xf transformA = input->transformA, transformB = input->transformB;
  Simplex simplex;
  simplex.readCache(cache, proxyA, transformA, proxyB, transformB);
  SimplexVertex* vertices = simplex.v;
  const int k_maxIters = 20;
  int saveA[3], saveB[3];
  int saveCount = 0;
  int iter = 0;
  while (iter < k_maxIters) {
    saveCount = simplex.count;
    for (int i = 0; i < saveCount; ++i) {
      saveA[i] = vertices[i].indexA;
      saveB[i] = vertices[i].indexB;
    }
    switch (simplex.count) {
      case 1:
        break;
      case 2:
        simplex.solve2();
        break;
      case 3:
        simplex.solve3();
        break;
    }
    if (simplex.count == 3) break;
    // (2.3.0 computes the closest point here but its "ensure progress" break is commented out.)
    V2 d = simplex.getSearchDirection();
    if (lengthSq(d) < kEps * kEps) break;
    SimplexVertex* vertex = vertices + simplex.count;
    vertex->indexA = proxyA->getSupport(mulT(transformA.q, -d));
    vertex->wA = mul(transformA, proxyA->vertices[vertex->indexA]);
    vertex->indexB = proxyB->getSupport(mulT(transformB.q, d));
    vertex->wB = mul(transformB, proxyB->vertices[vertex->indexB]);
    vertex->w = vertex->wB - vertex->wA;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < saveCount; ++i) {
      if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) {
        duplicate = true;
        break;
      }
    }
    if (duplicate) break;
    ++simplex.count;
  }
  simplex.getWitnessPoints(&output->pointA, &output->pointB);
  output->distance = distance(output->pointA, output->pointB);
  output->iterations = iter;
  simplex.writeCache(cache);
  if (input->useRadii) {
    float rA = proxyA->radius, rB = proxyB->radius;
    if (output->distance > rA + rB && output->distance > kEps) {
      output->distance -= rA + rB;
      V2 normal = output->pointB - output->pointA;
      normalize(normal);
      output->pointA += rA * normal;
      output->pointB -= rB * normal;
    } else {
      V2 p = 0.5f * (output->pointA + output->pointB);
      output->pointA = p;
      output->pointB = p;
      output->distance = 0.0f;
    }
  }
}

// b2TestOverlap(shapeA, shapeB, xfA, xfB) from b2Collision.cpp: the sensor test of b2Contact::Update.
bool testOverlapShapes(const Shape* a, const Shape* b, const Xf& xfA, const Xf& xfB) {
  DistanceInput input;
  input.proxyA.set(a);
  input.proxyB.set(b);
  input.transformA = xfA;
  input.transformB = xfB;
  input.useRadii = true;
  SimplexCache cache;
  cache.count = 0;
  DistanceOutput output;
  distanceGJK(&output, &cache, &input);
  return output.distance < 10.0f * kEps;
}

// =============================== b2TimeOfImpact.cpp ============================================
enum { SEP_POINTS = 0, SEP_FACE_A = 1, SEP_FACE_B = 2 };
struct SeparationFunction {
  const DistanceProxy* proxyA;
  const DistanceProxy* proxyB;
  Sweep sweepA, sweepB;
  int type;
  V2 localPoint, axis;

  float initialize(const SimplexCache* cache, const DistanceProxy* pA, const Sweep& sA, const DistanceProxy* pB,
                   const Sweep& sB, float t1) {
    proxyA = pA;
    proxyB = pB;
    int count = cache->count;
    sweepA = sA;
    sweepB = sB;
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t1);
    sweepB.getTransform(&xfB, t1);
    if (count == 1) {
      type = SEP_POINTS;
      V2 localPointA = proxyA->vertices[cache->indexA[0]];
      V2 localPointB = proxyB->vertices[cache->indexB[0]];
      V2 pointA = mul(xfA, localPointA);
      V2 pointB = mul(xfB, localPointB);
      axis = pointB - pointA;
      float s = normalize(axis);
      return s;
    } else if (cache->indexA[0] == cache->indexA[1]) {
      type = SEP_FACE_B;
      V2 localPointB1 = proxyB->vertices[cache->indexB[0]];
      V2 localPointB2 = proxyB->vertices[cache->indexB[1]];
      axis = cross(localPointB2 - localPointB1, 1.0f);
      normalize(axis);
      V2 normal = mul(xfB.q, axis);
      localPoint = 0.5f * (localPointB1 + localPointB2);
      V2 pointB = mul(xfB, localPoint);
      V2 localPointA = proxyA->vertices[cache->indexA[0]];
      V2 pointA = mul(xfA, localPointA);
      float s = dot(pointA - pointB, normal);
      if (s < 0.0f) {
        axis = -axis;
        s = -s;
      }
      return s;
    } else {
      type = SEP_FACE_A;
      V2 localPointA1 = proxyA->vertices[cache->indexA[0]];
      V2 localPointA2 = proxyA->vertices[cache->indexA[1]];
      axis = cross(localPointA2 - localPointA1, 1.0f);
      normalize(axis);
      V2 normal = mul(xfA.q, axis);
      localPoint = 0.5f * (localPointA1 + localPointA2);
      V2 pointA = mul(xfA, localPoint);
      V2 localPointB = proxyB->vertices[cache->indexB[0]];
      V2 pointB = mul(xfB, localPointB);
      float s = dot(pointB - pointA, normal);
      if (s < 0.0f) {
        axis = -axis;
        s = -s;
      }
      return s;
    }
  }
  float findMinSeparation(int* indexA, int* indexB, float t) const {
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t);
    sweepB.getTransform(&xfB, t);
    switch (type) {
      case SEP_POINTS: {
        V2 axisA = mulT(xfA.q, axis);
        V2 axisB = mulT(xfB.q, -axis);
        *indexA = proxyA->getSupport(axisA);
        *indexB = proxyB->getSupport(axisB);
        V2 pointA = mul(xfA, proxyA->vertices[*indexA]);
        V2 pointB = mul(xfB, proxyB->vertices[*indexB]);
        return dot(pointB - pointA, axis);
      }
      case SEP_FACE_A: {
        V2 normal = mul(xfA.q, axis);
        V2 pointA = mul(xfA, localPoint);
        V2 axisB = mulT(xfB.q, -normal);
        *indexA = -1;
        *indexB = proxyB->getSupport(axisB);
        V2 pointB = mul(xfB, proxyB->vertices[*indexB]);
        return dot(pointB - pointA, normal);
      }
      default: {
        V2 normal = mul(xfB.q, axis);
        V2 pointB = mul(xfB, localPoint);
        V2 axisA = mulT(xfA.q, -normal);
        *indexB = -1;
        *indexA = proxyA->getSupport(axisA);
        V2 pointA = mul(xfA, proxyA->vertices[*indexA]);
        return dot(pointA - pointB, normal);
      }
    }
  }
  float evaluate(int indexA, int indexB, float t) const {
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t);
    sweepB.getTransform(&xfB, t);
    switch (type) {
      case SEP_POINTS: {
        V2 pointA = mul(xfA, proxyA->vertices[indexA]);
        V2 pointB = mul(xfB, proxyB->vertices[indexB]);
        return dot(pointB - pointA, axis);
      }
      case SEP_FACE_A: {
        V2 normal = mul(xfA.q, axis);
        V2 pointA = mul(xfA, localPoint);
        V2 pointB = mul(xfB, proxyB->vertices[indexB]);
        return dot(pointB - pointA, normal);
      }
      default: {
        V2 normal = mul(xfB.q, axis);
        V2 pointB = mul(xfB, localPoint);
        V2 pointA = mul(xfA, proxyA->vertices[indexA]);
        return dot(pointA - pointB, normal);
      }
    }
  }
};

void timeOfImpact(TOIOutput* output, const TOIInput* input) {
  output->state = TOI_UNKNOWN;
  output->t = input->tMax;
  const DistanceProxy* proxyA = &input->proxyA;
  const DistanceProxy* proxyB = &input->proxyB;
  Sweep sweepA = input->sweepA, sweepB = input->sweepB;
  sweepA.normalizeAngles();
  sweepB.normalizeAngles();
  float tMax = input->tMax;
  float totalRadius = proxyA->radius + proxyB->radius;
  float target = fmax2(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
  float tolerance = 0.25f * kLinearSlop;
  float t1 = 0.0f;
  const int k_maxIterations = 20;
  int iter = 0;
  SimplexCache cache;
  cache.count = 0;
  DistanceInput distanceInput;
  distanceInput.proxyA = input->proxyA;
  distanceInput.proxyB = input->proxyB;
  distanceInput.useRadii = false;
  for (;;) {
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t1);
    sweepB.getTransform(&xfB, t1);
    distanceInput.transformA = xfA;
    distanceInput.transformB = xfB;
    DistanceOutput distanceOutput;
    distanceGJK(&distanceOutput, &cache, &distanceInput);
    if (distanceOutput.distance <= 0.0f) {
      output->state = TOI_OVERLAPPED;
      output->t = 0.0f;
      break;
    }
    if (distanceOutput.distance < target + tolerance) {
      output->state = TOI_TOUCHING;
      output->t = t1;
      break;
    }
    SeparationFunction fcn;
    fcn.initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1);
    bool done = false;
    float t2 = tMax;
    int pushBackIter = 0;
    for (;;) {
      int indexA, indexB;
      float s2 = fcn.findMinSeparation(&indexA, &indexB, t2);
      if (s2 > target + tolerance) {
        output->state = TOI_SEPARATED;
        output->t = tMax;
        done = true;
        break;
      }
      if (s2 > target - tolerance) {
        t1 = t2;
        break;
      }
      float s1 = fcn.evaluate(indexA, indexB, t1);
      if (s1 < target - tolerance) {
        output->state = TOI_FAILED;
        output->t = t1;
        done = true;
        break;
      }
      if (s1 <= target + tolerance) {
        output->state = TOI_TOUCHING;
        output->t = t1;
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
        float s = fcn.evaluate(indexA, indexB, t);
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
      if (pushBackIter == kMaxPolygonVertices) break;
    }
    ++iter;
    if (done) break;
    if (iter == k_maxIterations) {
      output->state = TOI_FAILED;
      output->t = t1;
      break;
    }
  }
}

// =============================== b2ContactSolver.cpp ===========================================
struct VelocityConstraintPoint {
  V2 rA, rB;
  float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias;
};
struct VelocityConstraint {
  VelocityConstraintPoint points[2];
  V2 normal;
  Mat22 normalMass, K;
  int indexA, indexB;
  float invMassA, invMassB, invIA, invIB;
  float friction, restitution;
  int pointCount, contactIndex;
};
struct PositionConstraint {
  V2 localPoints[2];
  V2 localNormal, localPoint;
  int indexA, indexB;
  float invMassA, invMassB;
  V2 localCenterA, localCenterB;
  float invIA, invIB;
  int type;
  float radiusA, radiusB;
  int pointCount;
};
struct Position {
  V2 c;
  float a;
};
struct Velocity {
  V2 v;
  float w;
};

struct WorldManifold {
  V2 normal;
  V2 points[2];
  void initialize(const Manifold* manifold, const Xf& xfA, float radiusA, const Xf& xfB, float radiusB) {
    if (manifold->pointCount == 0) return;
    switch (manifold->type) {
      case MANIFOLD_FACE_A: {
        normal = mul(xfA.q, manifold->localNormal);
        V2 planePoint = mul(xfA, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
          V2 clipPoint = mul(xfB, manifold->points[i].localPoint);
          V2 cA = clipPoint + (radiusA - dot(clipPoint - planePoint, normal)) * normal;
          V2 cB = clipPoint - radiusB * normal;
          points[i] = 0.5f * (cA + cB);
        }
      } break;
      case MANIFOLD_FACE_B: {
        normal = mul(xfB.q, manifold->localNormal);
        V2 planePoint = mul(xfB, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
          V2 clipPoint = mul(xfA, manifold->points[i].localPoint);
          V2 cB = clipPoint + (radiusB - dot(clipPoint - planePoint, normal)) * normal;
          V2 cA = clipPoint - radiusA * normal;
          points[i] = 0.5f * (cA + cB);
        }
        normal = -normal;
      } break;
      default:
        break;
    }
  }
};

struct ContactSolver {
  World* world;
  std::vector<Contact*>* contacts;
  std::vector<Position>* positions;
  std::vector<Velocity>* velocities;
  std::vector<VelocityConstraint> vcs;
  std::vector<PositionConstraint> pcs;

  ContactSolver(World* w, std::vector<Contact*>* cs, std::vector<Position>* ps, std::vector<Velocity>* vs,
                bool warmStarting, float dtRatio)
      : world(w), contacts(cs), positions(ps), velocities(vs) {
    int count = (int)cs->size();
    vcs.resize(count);
    pcs.resize(count);
    for (int i = 0; i < count; ++i) {
      Contact* contact = (*cs)[i];
      const Fixture& fixtureA = w->fixtures[contact->fA];
      const Fixture& fixtureB = w->fixtures[contact->fB];
      float radiusA = fixtureA.shape.radius, radiusB = fixtureB.shape.radius;
      const Body& bodyA = w->bodies[fixtureA.body];
      const Body& bodyB = w->bodies[fixtureB.body];
      const Manifold* manifold = &contact->manifold;
      int pointCount = manifold->pointCount;
      VelocityConstraint* vc = &vcs[i];
      vc->friction = contact->friction;
      vc->restitution = contact->restitution;
      vc->indexA = bodyA.islandIndex;
      vc->indexB = bodyB.islandIndex;
      vc->invMassA = bodyA.invMass;
      vc->invMassB = bodyB.invMass;
      vc->invIA = bodyA.invI;
      vc->invIB = bodyB.invI;
      vc->contactIndex = i;
      vc->pointCount = pointCount;
      vc->K.ex = vc->K.ey = mk(0, 0);
      vc->normalMass.ex = vc->normalMass.ey = mk(0, 0);
      PositionConstraint* pc = &pcs[i];
      pc->indexA = bodyA.islandIndex;
      pc->indexB = bodyB.islandIndex;
      pc->invMassA = bodyA.invMass;
      pc->invMassB = bodyB.invMass;
      pc->localCenterA = bodyA.sweep.localCenter;
      pc->localCenterB = bodyB.sweep.localCenter;
      pc->invIA = bodyA.invI;
      pc->invIB = bodyB.invI;
      pc->localNormal = manifold->localNormal;
      pc->localPoint = manifold->localPoint;
      pc->pointCount = pointCount;
      pc->radiusA = radiusA;
      pc->radiusB = radiusB;
      pc->type = manifold->type;
      for (int j = 0; j < pointCount; ++j) {
        const ManifoldPoint* cp = manifold->points + j;
        VelocityConstraintPoint* vcp = vc->points + j;
        if (warmStarting) {
          vcp->normalImpulse = dtRatio * cp->normalImpulse;
          vcp->tangentImpulse = dtRatio * cp->tangentImpulse;
        } else {
          vcp->normalImpulse = 0.0f;
          vcp->tangentImpulse = 0.0f;
        }
        vcp->rA = vcp->rB = mk(0, 0);
        vcp->normalMass = vcp->tangentMass = vcp->velocityBias = 0.0f;
        pc->localPoints[j] = cp->localPoint;
      }
    }
  }

  void initializeVelocityConstraints() {
    for (size_t i = 0; i < vcs.size(); ++i) {
      VelocityConstraint* vc = &vcs[i];
      PositionConstraint* pc = &pcs[i];
      float radiusA = pc->radiusA, radiusB = pc->radiusB;
      const Manifold* manifold = &(*contacts)[vc->contactIndex]->manifold;
      int indexA = vc->indexA, indexB = vc->indexB;
      float mA = vc->invMassA, mB = vc->invMassB, iA = vc->invIA, iB = vc->invIB;
      V2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
      V2 cA = (*positions)[indexA].c;
      float aA = (*positions)[indexA].a;
      V2 vA = (*velocities)[indexA].v;
      float wA = (*velocities)[indexA].w;
      V2 cB = (*positions)[indexB].c;
      float aB = (*positions)[indexB].a;
      V2 vB = (*velocities)[indexB].v;
      float wB = (*velocities)[indexB].w;
      Xf xfA, xfB;
      xfA.q.set(aA);
      xfB.q.set(aB);
      xfA.p = cA - mul(xfA.q, localCenterA);
      xfB.p = cB - mul(xfB.q, localCenterB);
      WorldManifold wm;
      wm.initialize(manifold, xfA, radiusA, xfB, radiusB);
      vc->normal = wm.normal;
      int pointCount = vc->pointCount;
      for (int j = 0; j < pointCount; ++j) {
        VelocityConstraintPoint* vcp = vc->points + j;
        vcp->rA = wm.points[j] - cA;
        vcp->rB = wm.points[j] - cB;
        float rnA = cross(vcp->rA, vc->normal);
        float rnB = cross(vcp->rB, vc->normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        vcp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
        V2 tangent = cross(vc->normal, 1.0f);
        float rtA = cross(vcp->rA, tangent);
        float rtB = cross(vcp->rB, tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        vcp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
        vcp->velocityBias = 0.0f;
        float vRel = dot(vc->normal, vB + cross(wB, vcp->rB) - vA - cross(wA, vcp->rA));
        if (vRel < -kVelocityThreshold) vcp->velocityBias = -vc->restitution * vRel;
      }
      if (vc->pointCount == 2) {
        VelocityConstraintPoint* vcp1 = vc->points + 0;
        VelocityConstraintPoint* vcp2 = vc->points + 1;
        float rn1A = cross(vcp1->rA, vc->normal);
        float rn1B = cross(vcp1->rB, vc->normal);
        float rn2A = cross(vcp2->rA, vc->normal);
        float rn2B = cross(vcp2->rB, vc->normal);
        float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_maxConditionNumber = 1000.0f;
        if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
          vc->K.ex = mk(k11, k12);
          vc->K.ey = mk(k12, k22);
          vc->normalMass = vc->K.inverse();
        } else {
          vc->pointCount = 1;
        }
      }
    }
  }

  void warmStart() {
    for (size_t i = 0; i < vcs.size(); ++i) {
      VelocityConstraint* vc = &vcs[i];
      int indexA = vc->indexA, indexB = vc->indexB;
      float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
      int pointCount = vc->pointCount;
      V2 vA = (*velocities)[indexA].v;
      float wA = (*velocities)[indexA].w;
      V2 vB = (*velocities)[indexB].v;
      float wB = (*velocities)[indexB].w;
      V2 normal = vc->normal;
      V2 tangent = cross(normal, 1.0f);
      for (int j = 0; j < pointCount; ++j) {
        VelocityConstraintPoint* vcp = vc->points + j;
        V2 P = vcp->normalImpulse * normal + vcp->tangentImpulse * tangent;
        wA -= iA * cross(vcp->rA, P);
        vA -= mA * P;
        wB += iB * cross(vcp->rB, P);
        vB += mB * P;
      }
      (*velocities)[indexA].v = vA;
      (*velocities)[indexA].w = wA;
      (*velocities)[indexB].v = vB;
      (*velocities)[indexB].w = wB;
    }
  }

  void solveVelocityConstraints() {
    for (size_t i = 0; i < vcs.size(); ++i) {
      VelocityConstraint* vc = &vcs[i];
      int indexA = vc->indexA, indexB = vc->indexB;
      float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
      int pointCount = vc->pointCount;
      V2 vA = (*velocities)[indexA].v;
      float wA = (*velocities)[indexA].w;
      V2 vB = (*velocities)[indexB].v;
      float wB = (*velocities)[indexB].w;
      V2 normal = vc->normal;
      V2 tangent = cross(normal, 1.0f);
      float friction = vc->friction;
      for (int j = 0; j < pointCount; ++j) {
        VelocityConstraintPoint* vcp = vc->points + j;
        V2 dv = vB + cross(wB, vcp->rB) - vA - cross(wA, vcp->rA);
        float vt = dot(dv, tangent) - 0.0f;  // tangentSpeed == 0
        float lambda = vcp->tangentMass * (-vt);
        float maxFriction = friction * vcp->normalImpulse;
        float newImpulse = fclamp(vcp->tangentImpulse + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - vcp->tangentImpulse;
        vcp->tangentImpulse = newImpulse;
        V2 P = lambda * tangent;
        vA -= mA * P;
        wA -= iA * cross(vcp->rA, P);
        vB += mB * P;
        wB += iB * cross(vcp->rB, P);
      }
      if (vc->pointCount == 1) {
        VelocityConstraintPoint* vcp = vc->points + 0;
        V2 dv = vB + cross(wB, vcp->rB) - vA - cross(wA, vcp->rA);
        float vn = dot(dv, normal);
        float lambda = -vcp->normalMass * (vn - vcp->velocityBias);
        float newImpulse = fmax2(vcp->normalImpulse + lambda, 0.0f);
        lambda = newImpulse - vcp->normalImpulse;
        vcp->normalImpulse = newImpulse;
        V2 P = lambda * normal;
        vA -= mA * P;
        wA -= iA * cross(vcp->rA, P);
        vB += mB * P;
        wB += iB * cross(vcp->rB, P);
      } else {
        VelocityConstraintPoint* cp1 = vc->points + 0;
        VelocityConstraintPoint* cp2 = vc->points + 1;
        V2 a = mk(cp1->normalImpulse, cp2->normalImpulse);
        V2 dv1 = vB + cross(wB, cp1->rB) - vA - cross(wA, cp1->rA);
        V2 dv2 = vB + cross(wB, cp2->rB) - vA - cross(wA, cp2->rA);
        float vn1 = dot(dv1, normal);
        float vn2 = dot(dv2, normal);
        V2 b;
        b.x = vn1 - cp1->velocityBias;
        b.y = vn2 - cp2->velocityBias;
        b -= mul(vc->K, a);
        for (;;) {
          V2 x = -mul(vc->normalMass, b);
          if (x.x >= 0.0f && x.y >= 0.0f) {
            V2 d = x - a;
            V2 P1 = d.x * normal, P2 = d.y * normal;
            vA -= mA * (P1 + P2);
            wA -= iA * (cross(cp1->rA, P1) + cross(cp2->rA, P2));
            vB += mB * (P1 + P2);
            wB += iB * (cross(cp1->rB, P1) + cross(cp2->rB, P2));
            cp1->normalImpulse = x.x;
            cp2->normalImpulse = x.y;
            break;
          }
          x.x = -cp1->normalMass * b.x;
          x.y = 0.0f;
          vn1 = 0.0f;
          vn2 = vc->K.ex.y * x.x + b.y;
          if (x.x >= 0.0f && vn2 >= 0.0f) {
            V2 d = x - a;
            V2 P1 = d.x * normal, P2 = d.y * normal;
            vA -= mA * (P1 + P2);
            wA -= iA * (cross(cp1->rA, P1) + cross(cp2->rA, P2));
            vB += mB * (P1 + P2);
            wB += iB * (cross(cp1->rB, P1) + cross(cp2->rB, P2));
            cp1->normalImpulse = x.x;
            cp2->normalImpulse = x.y;
            break;
          }
          x.x = 0.0f;
          x.y = -cp2->normalMass * b.y;
          vn1 = vc->K.ey.x * x.y + b.x;
          vn2 = 0.0f;
          if (x.y >= 0.0f && vn1 >= 0.0f) {
            V2 d = x - a;
            V2 P1 = d.x * normal, P2 = d.y * normal;
            vA -= mA * (P1 + P2);
            wA -= iA * (cross(cp1->rA, P1) + cross(cp2->rA, P2));
            vB += mB * (P1 + P2);
            wB += iB * (cross(cp1->rB, P1) + cross(cp2->rB, P2));
            cp1->normalImpulse = x.x;
            cp2->normalImpulse = x.y;
            break;
          }
          x.x = 0.0f;
          x.y = 0.0f;
          vn1 = b.x;
          vn2 = b.y;
          if (vn1 >= 0.0f && vn2 >= 0.0f) {
            V2 d = x - a;
            V2 P1 = d.x * normal, P2 = d.y * normal;
            vA -= mA * (P1 + P2);
            wA -= iA * (cross(cp1->rA, P1) + cross(cp2->rA, P2));
            vB += mB * (P1 + P2);
            wB += iB * (cross(cp1->rB, P1) + cross(cp2->rB, P2));
            cp1->normalImpulse = x.x;
            cp2->normalImpulse = x.y;
            break;
          }
          break;
        }
      }
      (*velocities)[indexA].v = vA;
      (*velocities)[indexA].w = wA;
      (*velocities)[indexB].v = vB;
      (*velocities)[indexB].w = wB;
    }
  }

  void storeImpulses() {
    for (size_t i = 0; i < vcs.size(); ++i) {
      VelocityConstraint* vc = &vcs[i];
      Manifold* manifold = &(*contacts)[vc->contactIndex]->manifold;
      for (int j = 0; j < vc->pointCount; ++j) {
        manifold->points[j].normalImpulse = vc->points[j].normalImpulse;
        manifold->points[j].tangentImpulse = vc->points[j].tangentImpulse;
      }
    }
  }

  static void psmInitialize(const PositionConstraint* pc, const Xf& xfA, const Xf& xfB, int index, V2* normal,
                            V2* point, float* separation) {
    switch (pc->type) {
      case MANIFOLD_FACE_A: {
        *normal = mul(xfA.q, pc->localNormal);
        V2 planePoint = mul(xfA, pc->localPoint);
        V2 clipPoint = mul(xfB, pc->localPoints[index]);
        *separation = dot(clipPoint - planePoint, *normal) - pc->radiusA - pc->radiusB;
        *point = clipPoint;
      } break;
      default: {
        *normal = mul(xfB.q, pc->localNormal);
        V2 planePoint = mul(xfB, pc->localPoint);
        V2 clipPoint = mul(xfA, pc->localPoints[index]);
        *separation = dot(clipPoint - planePoint, *normal) - pc->radiusA - pc->radiusB;
        *point = clipPoint;
        *normal = -*normal;
      } break;
    }
  }

  // toi == false: b2ContactSolver::SolvePositionConstraints; toi == true: SolveTOIPositionConstraints
  bool solvePositionConstraints(bool toi, int toiIndexA, int toiIndexB) {
    float minSeparation = 0.0f;
    for (size_t i = 0; i < pcs.size(); ++i) {
      PositionConstraint* pc = &pcs[i];
      int indexA = pc->indexA, indexB = pc->indexB;
      V2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
      int pointCount = pc->pointCount;
      float mA, iA, mB, iB;
      if (toi) {
        mA = 0.0f;
        iA = 0.0f;
        if (indexA == toiIndexA || indexA == toiIndexB) {
          mA = pc->invMassA;
          iA = pc->invIA;
        }
        mB = 0.0f;
        iB = 0.0f;
        if (indexB == toiIndexA || indexB == toiIndexB) {
          mB = pc->invMassB;
          iB = pc->invIB;
        }
      } else {
        mA = pc->invMassA;
        iA = pc->invIA;
        mB = pc->invMassB;
        iB = pc->invIB;
      }
      V2 cA = (*positions)[indexA].c;
      float aA = (*positions)[indexA].a;
      V2 cB = (*positions)[indexB].c;
      float aB = (*positions)[indexB].a;
      for (int j = 0; j < pointCount; ++j) {
        Xf xfA, xfB;
        xfA.q.set(aA);
        xfB.q.set(aB);
        xfA.p = cA - mul(xfA.q, localCenterA);
        xfB.p = cB - mul(xfB.q, localCenterB);
        V2 normal, point;
        float separation;
        psmInitialize(pc, xfA, xfB, j, &normal, &point, &separation);
        V2 rA = point - cA;
        V2 rB = point - cB;
        minSeparation = fmin2(minSeparation, separation);
        float C = fclamp((toi ? kToiBaumgarte : kBaumgarte) * (separation + kLinearSlop), -kMaxLinearCorrection,
                         0.0f);
        float rnA = cross(rA, normal);
        float rnB = cross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        V2 P = impulse * normal;
        cA -= mA * P;
        aA -= iA * cross(rA, P);
        cB += mB * P;
        aB += iB * cross(rB, P);
      }
      (*positions)[indexA].c = cA;
      (*positions)[indexA].a = aA;
      (*positions)[indexB].c = cB;
      (*positions)[indexB].a = aB;
    }
    return minSeparation >= (toi ? -1.5f * kLinearSlop : -3.0f * kLinearSlop);
  }
};

// =============================== b2World / b2ContactManager / b2Island ==========================
void World::clear() {
  for (Contact* c : contacts) delete c;
  contacts.clear();
  bodies.clear();
  fixtures.clear();
  moveBuffer.clear();
  newFixture = false;
}

int World::createBody(int type, V2 position, float angle) {
  Body b;
  std::memset(&b, 0, sizeof(b));
  b.type = type;
  b.xf.p = position;
  b.xf.q.set(angle);
  b.sweep.localCenter = mk(0, 0);
  b.sweep.c0 = b.sweep.c = position;
  b.sweep.a0 = b.sweep.a = angle;
  b.sweep.alpha0 = 0.0f;
  b.v = mk(0, 0);
  b.w = 0.0f;
  b.force = mk(0, 0);
  b.torque = 0.0f;
  b.linearDamping = 0.0f;
  b.angularDamping = 0.0f;
  b.awake = true;
  b.islandFlag = false;
  b.sleepTime = 0.0f;
  if (type == BODY_DYNAMIC) {
    b.mass = 1.0f;
    b.invMass = 1.0f;
  } else {
    b.mass = 0.0f;
    b.invMass = 0.0f;
  }
  b.I = 0.0f;
  b.invI = 0.0f;
  b.fixtureBegin = b.fixtureEnd = (int)fixtures.size();
  bodies.push_back(b);
  return (int)bodies.size() - 1;
}

// b2Body::CreateFixture + b2Fixture::CreateProxies + b2Body::ResetMassData (single-fixture bodies
// and zero-density multi-fixture statics are all this scene has).
int World::createFixture(int body, const Shape& shape, float density, float friction, float restitution,
                         unsigned cat, unsigned mask, bool sensor) {
  Fixture f;
  f.body = body;
  f.shape = shape;
  f.density = density;
  f.friction = friction;
  f.restitution = restitution;
  f.categoryBits = cat;
  f.maskBits = mask;
  f.isSensor = sensor;
  Body& b = bodies[body];
  AABB aabb;
  f.shape.computeAABB(&aabb, b.xf);
  V2 r = mk(kAabbExtension, kAabbExtension);
  f.fatAABB.lo = aabb.lo - r;
  f.fatAABB.hi = aabb.hi + r;
  fixtures.push_back(f);
  int fi = (int)fixtures.size() - 1;
  b.fixtureEnd = fi + 1;
  moveBuffer.push_back(fi);
  newFixture = true;
  if (density > 0.0f && b.type == BODY_DYNAMIC) {
    // ResetMassData
    b.mass = 0.0f;
    b.invMass = 0.0f;
    b.I = 0.0f;
    b.invI = 0.0f;
    V2 localCenter = mk(0, 0);
    MassData md;
    fixtures[fi].shape.computeMass(&md, density);
    b.mass += md.mass;
    localCenter += md.mass * md.center;
    b.I += md.I;
    if (b.mass > 0.0f) {
      b.invMass = 1.0f / b.mass;
      localCenter *= b.invMass;
    } else {
      b.mass = 1.0f;
      b.invMass = 1.0f;
    }
    if (b.I > 0.0f) {
      b.I -= b.mass * dot(localCenter, localCenter);
      b.invI = 1.0f / b.I;
    } else {
      b.I = 0.0f;
      b.invI = 0.0f;
    }
    V2 oldCenter = b.sweep.c;
    b.sweep.localCenter = localCenter;
    b.sweep.c0 = b.sweep.c = mul(b.xf, b.sweep.localCenter);
    b.v += cross(b.w, b.sweep.c - oldCenter);
  }
  return fi;
}

void World::moveProxy(int fi, const AABB& aabb, V2 displacement) {
  Fixture& f = fixtures[fi];
  if (f.fatAABB.contains(aabb)) return;
  AABB b = aabb;
  V2 r = mk(kAabbExtension, kAabbExtension);
  b.lo = b.lo - r;
  b.hi = b.hi + r;
  V2 d = kAabbMultiplier * displacement;
  if (d.x < 0.0f)
    b.lo.x += d.x;
  else
    b.hi.x += d.x;
  if (d.y < 0.0f)
    b.lo.y += d.y;
  else
    b.hi.y += d.y;
  f.fatAABB = b;
  // b2BroadPhase::BufferMove
  moveBuffer.push_back(fi);
}

// b2Body::SetTransform (2.3.0: no FindNewContacts here)
void World::setTransform(int bi, V2 position, float angle) {
  Body& b = bodies[bi];
  b.xf.q.set(angle);
  b.xf.p = position;
  b.sweep.c = mul(b.xf, b.sweep.localCenter);
  b.sweep.a = angle;
  b.sweep.c0 = b.sweep.c;
  b.sweep.a0 = angle;
  for (int fi = b.fixtureBegin; fi < b.fixtureEnd; ++fi) {
    AABB a1;
    fixtures[fi].shape.computeAABB(&a1, b.xf);
    AABB comb;
    comb.combine(a1, a1);
    moveProxy(fi, comb, b.xf.p - b.xf.p);
  }
}

void World::synchronizeFixtures(int bi) {
  Body& b = bodies[bi];
  Xf xf1;
  xf1.q.set(b.sweep.a0);
  xf1.p = b.sweep.c0 - mul(xf1.q, b.sweep.localCenter);
  for (int fi = b.fixtureBegin; fi < b.fixtureEnd; ++fi) {
    AABB a1, a2, comb;
    fixtures[fi].shape.computeAABB(&a1, xf1);
    fixtures[fi].shape.computeAABB(&a2, b.xf);
    comb.combine(a1, a2);
    V2 displacement = b.xf.p - xf1.p;
    moveProxy(fi, comb, displacement);
  }
}

Contact* World::findContact(int fA, int fB) {
  for (Contact* c : contacts)
    if ((c->fA == fA && c->fB == fB) || (c->fA == fB && c->fB == fA)) return c;
  return nullptr;
}

// b2ContactManager::FindNewContacts -> b2BroadPhase::UpdatePairs -> b2ContactManager::AddPair
void World::findNewContacts() {
  std::vector<std::pair<int, int>> pairs;
  for (int q : moveBuffer) {
    const AABB& fat = fixtures[q].fatAABB;
    for (int p = 0; p < (int)fixtures.size(); ++p) {
      if (p == q) continue;
      if (!testOverlap(fixtures[p].fatAABB, fat)) continue;
      pairs.push_back(std::make_pair(std::min(p, q), std::max(p, q)));
    }
  }
  moveBuffer.clear();
  std::sort(pairs.begin(), pairs.end());
  pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
  for (auto& pr : pairs) {
    int fA = pr.first, fB = pr.second;
    Fixture& fixA = fixtures[fA];
    Fixture& fixB = fixtures[fB];
    if (fixA.body == fixB.body) continue;
    if (findContact(fA, fB)) continue;
    Body& bodyA = bodies[fixA.body];
    Body& bodyB = bodies[fixB.body];
    if (bodyA.type != BODY_DYNAMIC && bodyB.type != BODY_DYNAMIC) continue;  // b2Body::ShouldCollide
    bool collide = (fixA.maskBits & fixB.categoryBits) != 0 && (fixA.categoryBits & fixB.maskBits) != 0;
    if (!collide) continue;
    // b2Contact::Create: for polygon-vs-circle the polygon is always fixture A
    if (fixA.shape.type == SHAPE_CIRCLE && fixB.shape.type == SHAPE_POLYGON) std::swap(fA, fB);
    Contact* c = new Contact();
    c->fA = fA;
    c->fB = fB;
    c->manifold.pointCount = 0;
    c->touching = false;
    c->enabled = true;
    c->islandFlag = false;
    c->toiFlag = false;
    c->toiCount = 0;
    c->toi = 1.0f;
    c->friction = sqrtf(fixtures[fA].friction * fixtures[fB].friction);
    c->restitution = fixtures[fA].restitution > fixtures[fB].restitution ? fixtures[fA].restitution
                                                                            : fixtures[fB].restitution;
    contacts.insert(contacts.begin(), c);
    if (!fixtures[fA].isSensor && !fixtures[fB].isSensor) {
      bodies[fixtures[fA].body].setAwake(true);
      bodies[fixtures[fB].body].setAwake(true);
    }
  }
}

void World::destroyContact(size_t idx) {
  delete contacts[idx];
  contacts.erase(contacts.begin() + idx);
}

// b2Contact::Update
void World::updateContact(Contact* c) {
  Manifold oldManifold = c->manifold;
  c->enabled = true;
  bool touching = false;
  bool wasTouching = c->touching;
  Fixture& fixA = fixtures[c->fA];
  Fixture& fixB = fixtures[c->fB];
  bool sensor = fixA.isSensor || fixB.isSensor;
  Body& bodyA = bodies[fixA.body];
  Body& bodyB = bodies[fixB.body];
  const Xf& xfA = bodyA.xf;
  const Xf& xfB = bodyB.xf;
  if (sensor) {
    touching = testOverlapShapes(&fixA.shape, &fixB.shape, xfA, xfB);
    c->manifold.pointCount = 0;
  } else {
    if (fixB.shape.type == SHAPE_CIRCLE)
      collidePolygonAndCircle(&c->manifold, &fixA.shape, xfA, &fixB.shape, xfB);
    else
      collidePolygons(&c->manifold, &fixA.shape, xfA, &fixB.shape, xfB);
    touching = c->manifold.pointCount > 0;
    for (int i = 0; i < c->manifold.pointCount; ++i) {
      ManifoldPoint* mp2 = c->manifold.points + i;
      mp2->normalImpulse = 0.0f;
      mp2->tangentImpulse = 0.0f;
      uint32_t id2 = mp2->key;
      for (int j = 0; j < oldManifold.pointCount; ++j) {
        ManifoldPoint* mp1 = oldManifold.points + j;
        if (mp1->key == id2) {
          mp2->normalImpulse = mp1->normalImpulse;
          mp2->tangentImpulse = mp1->tangentImpulse;
          break;
        }
      }
    }
    if (touching != wasTouching) {
      bodyA.setAwake(true);
      bodyB.setAwake(true);
    }
  }
  c->touching = touching;
  if (!wasTouching && touching && beginContact) beginContact(listenerUser, c);
}

// b2ContactManager::Collide
void World::collide() {
  size_t i = 0;
  while (i < contacts.size()) {
    Contact* c = contacts[i];
    Body& bodyA = bodies[fixtures[c->fA].body];
    Body& bodyB = bodies[fixtures[c->fB].body];
    bool activeA = bodyA.awake && bodyA.type != BODY_STATIC;
    bool activeB = bodyB.awake && bodyB.type != BODY_STATIC;
    if (!activeA && !activeB) {
      ++i;
      continue;
    }
    bool overlap = testOverlap(fixtures[c->fA].fatAABB, fixtures[c->fB].fatAABB);
    if (!overlap) {
      destroyContact(i);
      continue;
    }
    updateContact(c);
    ++i;
  }
}

void World::clearForces() {
  for (Body& b : bodies) {
    b.force = mk(0, 0);
    b.torque = 0.0f;
  }
}

// b2World::Solve + b2Island::Solve
void World::solve(float h, int velIters, int posIters) {
  for (Body& b : bodies) b.islandFlag = false;
  for (Contact* c : contacts) c->islandFlag = false;
  std::vector<int> stack;
  std::vector<int> islandBodies;
  std::vector<Contact*> islandContacts;
  for (int seed = (int)bodies.size() - 1; seed >= 0; --seed) {  // body list is newest-first
    Body& sb = bodies[seed];
    if (sb.islandFlag) continue;
    if (!sb.awake) continue;
    if (sb.type == BODY_STATIC) continue;
    islandBodies.clear();
    islandContacts.clear();
    stack.clear();
    stack.push_back(seed);
    sb.islandFlag = true;
    while (!stack.empty()) {
      int bi = stack.back();
      stack.pop_back();
      Body& b = bodies[bi];
      b.islandIndex = (int)islandBodies.size();
      islandBodies.push_back(bi);
      b.setAwake(true);
      if (b.type == BODY_STATIC) continue;
      // the body's contact-edge list has the same relative order as the world contact list
      for (Contact* c : contacts) {
        int bA = fixtures[c->fA].body, bB = fixtures[c->fB].body;
        if (bA != bi && bB != bi) continue;
        if (c->islandFlag) continue;
        if (!c->enabled || !c->touching) continue;
        if (fixtures[c->fA].isSensor || fixtures[c->fB].isSensor) continue;
        islandContacts.push_back(c);
        c->islandFlag = true;
        int other = (bA == bi) ? bB : bA;
        if (bodies[other].islandFlag) continue;
        stack.push_back(other);
        bodies[other].islandFlag = true;
      }
    }

    // ---- b2Island::Solve ----
    int nb = (int)islandBodies.size();
    std::vector<Position> positions(nb);
    std::vector<Velocity> velocities(nb);
    for (int i = 0; i < nb; ++i) {
      Body& b = bodies[islandBodies[i]];
      V2 c = b.sweep.c;
      float a = b.sweep.a;
      V2 v = b.v;
      float w = b.w;
      b.sweep.c0 = b.sweep.c;
      b.sweep.a0 = b.sweep.a;
      if (b.type == BODY_DYNAMIC) {
        // gravity is (0,0) (hockey_env.py:107); gravityScale 1
        v += h * (1.0f * mk(0.0f, 0.0f) + b.invMass * b.force);
        w += h * b.invI * b.torque;
        // 2.3.0 clamp-form damping (confirmed by the notebook trace)
        v *= fclamp(1.0f - h * b.linearDamping, 0.0f, 1.0f);
        w *= fclamp(1.0f - h * b.angularDamping, 0.0f, 1.0f);
      }
      positions[i].c = c;
      positions[i].a = a;
      velocities[i].v = v;
      velocities[i].w = w;
    }
    const float dtRatio = 1.0f;  // inv_dt0 * dt, fixed step
    ContactSolver solver(this, &islandContacts, &positions, &velocities, true, dtRatio);
    solver.initializeVelocityConstraints();
    solver.warmStart();
    for (int i = 0; i < velIters; ++i) solver.solveVelocityConstraints();
    solver.storeImpulses();
    for (int i = 0; i < nb; ++i) {
      V2 c = positions[i].c;
      float a = positions[i].a;
      V2 v = velocities[i].v;
      float w = velocities[i].w;
      V2 translation = h * v;
      if (dot(translation, translation) > kMaxTranslationSquared) {
        float ratio = kMaxTranslation / length(translation);
        v *= ratio;
      }
      float rotation = h * w;
      if (rotation * rotation > kMaxRotationSquared) {
        float ratio = kMaxRotation / fabs2(rotation);
        w *= ratio;
      }
      c += h * v;
      a += h * w;
      positions[i].c = c;
      positions[i].a = a;
      velocities[i].v = v;
      velocities[i].w = w;
    }
    bool positionSolved = false;
    for (int i = 0; i < posIters; ++i) {
      bool contactsOkay = solver.solvePositionConstraints(false, 0, 0);
      if (contactsOkay) {
        positionSolved = true;
        break;
      }
    }
    for (int i = 0; i < nb; ++i) {
      Body& b = bodies[islandBodies[i]];
      b.sweep.c = positions[i].c;
      b.sweep.a = positions[i].a;
      b.v = velocities[i].v;
      b.w = velocities[i].w;
      b.synchronizeTransform();
    }
    {
      float minSleepTime = kMaxFloat;
      const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
      const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
      for (int i = 0; i < nb; ++i) {
        Body& b = bodies[islandBodies[i]];
        if (b.type == BODY_STATIC) continue;
        if (b.w * b.w > angTolSqr || dot(b.v, b.v) > linTolSqr) {
          b.sleepTime = 0.0f;
          minSleepTime = 0.0f;
        } else {
          b.sleepTime += h;
          minSleepTime = fmin2(minSleepTime, b.sleepTime);
        }
      }
      if (minSleepTime >= kTimeToSleep && positionSolved) {
        for (int i = 0; i < nb; ++i) bodies[islandBodies[i]].setAwake(false);
      }
    }
    for (int i = 0; i < nb; ++i) {
      Body& b = bodies[islandBodies[i]];
      if (b.type == BODY_STATIC) b.islandFlag = false;
    }
  }
  // Synchronize fixtures (body list order), look for new contacts.
  for (int bi = (int)bodies.size() - 1; bi >= 0; --bi) {
    Body& b = bodies[bi];
    if (!b.islandFlag) continue;
    if (b.type == BODY_STATIC) continue;
    synchronizeFixtures(bi);
  }
  findNewContacts();
}

// b2World::SolveTOI + b2Island::SolveTOI
void World::solveTOI(float dt, int velIters) {
  for (Body& b : bodies) {
    b.islandFlag = false;
    b.sweep.alpha0 = 0.0f;
  }
  for (Contact* c : contacts) {
    c->toiFlag = false;
    c->islandFlag = false;
    c->toiCount = 0;
    c->toi = 1.0f;
  }
  for (;;) {
    Contact* minContact = nullptr;
    float minAlpha = 1.0f;
    for (Contact* c : contacts) {
      if (!c->enabled) continue;
      if (c->toiCount > kMaxSubSteps) continue;
      float alpha = 1.0f;
      if (c->toiFlag) {
        alpha = c->toi;
      } else {
        Fixture& fA = fixtures[c->fA];
        Fixture& fB = fixtures[c->fB];
        if (fA.isSensor || fB.isSensor) continue;
        Body& bA = bodies[fA.body];
        Body& bB = bodies[fB.body];
        bool activeA = bA.awake && bA.type != BODY_STATIC;
        bool activeB = bB.awake && bB.type != BODY_STATIC;
        if (!activeA && !activeB) continue;
        bool collideA = bA.type != BODY_DYNAMIC;  // no bullets in this scene
        bool collideB = bB.type != BODY_DYNAMIC;
        if (!collideA && !collideB) continue;
        float alpha0 = bA.sweep.alpha0;
        if (bA.sweep.alpha0 < bB.sweep.alpha0) {
          alpha0 = bB.sweep.alpha0;
          if (bA.type == BODY_STATIC && !g_static_drift) bA.sweep.alpha0 = alpha0; else bA.sweep.advance(alpha0);
        } else if (bB.sweep.alpha0 < bA.sweep.alpha0) {
          alpha0 = bA.sweep.alpha0;
          if (bB.type == BODY_STATIC && !g_static_drift) bB.sweep.alpha0 = alpha0; else bB.sweep.advance(alpha0);
        }
        TOIInput input;
        input.proxyA.set(&fA.shape);
        input.proxyB.set(&fB.shape);
        input.sweepA = bA.sweep;
        input.sweepB = bB.sweep;
        input.tMax = 1.0f;
        TOIOutput output;
        timeOfImpact(&output, &input);
        ++nToiCalls;
        float beta = output.t;
        if (output.state == TOI_TOUCHING)
          alpha = fmin2(alpha0 + (1.0f - alpha0) * beta, 1.0f);
        else
          alpha = 1.0f;
        c->toi = alpha;
        c->toiFlag = true;
      }
      if (alpha < minAlpha) {
        minContact = c;
        minAlpha = alpha;
      }
    }
    if (minContact == nullptr || 1.0f - 10.0f * kEps < minAlpha) break;
    ++nToiEvents;

    Fixture& fA = fixtures[minContact->fA];
    Fixture& fB = fixtures[minContact->fB];
    int biA = fA.body, biB = fB.body;
    Body& bA = bodies[biA];
    Body& bB = bodies[biB];
    Sweep backup1 = bA.sweep, backup2 = bB.sweep;
    bA.advance(minAlpha);
    bB.advance(minAlpha);
    updateContact(minContact);
    minContact->toiFlag = false;
    ++minContact->toiCount;
    if (!minContact->enabled || !minContact->touching) {
      minContact->enabled = false;
      bA.sweep = backup1;
      bB.sweep = backup2;
      bA.synchronizeTransform();
      bB.synchronizeTransform();
      continue;
    }
    bA.setAwake(true);
    bB.setAwake(true);
    std::vector<int> islandBodies;
    std::vector<Contact*> islandContacts;
    bA.islandIndex = 0;
    islandBodies.push_back(biA);
    bB.islandIndex = 1;
    islandBodies.push_back(biB);
    islandContacts.push_back(minContact);
    bA.islandFlag = true;
    bB.islandFlag = true;
    minContact->islandFlag = true;
    int pair[2] = {biA, biB};
    for (int k = 0; k < 2; ++k) {
      int bi = pair[k];
      Body& body = bodies[bi];
      if (body.type != BODY_DYNAMIC) continue;
      for (Contact* contact : contacts) {
        int cA = fixtures[contact->fA].body, cB = fixtures[contact->fB].body;
        if (cA != bi && cB != bi) continue;
        if ((int)islandBodies.size() == 2 * kMaxTOIContacts) break;
        if ((int)islandContacts.size() == kMaxTOIContacts) break;
        if (contact->islandFlag) continue;
        int oi = (cA == bi) ? cB : cA;
        Body& other = bodies[oi];
        if (other.type == BODY_DYNAMIC) continue;  // no bullets
        if (fixtures[contact->fA].isSensor || fixtures[contact->fB].isSensor) continue;
        Sweep backup = other.sweep;
        if (!other.islandFlag) other.advance(minAlpha);
        updateContact(contact);
        if (!contact->enabled) {
          other.sweep = backup;
          other.synchronizeTransform();
          continue;
        }
        if (!contact->touching) {
          other.sweep = backup;
          other.synchronizeTransform();
          continue;
        }
        contact->islandFlag = true;
        islandContacts.push_back(contact);
        if (other.islandFlag) continue;
        other.islandFlag = true;
        if (other.type != BODY_STATIC) other.setAwake(true);
        other.islandIndex = (int)islandBodies.size();
        islandBodies.push_back(oi);
      }
    }
    float subDt = (1.0f - minAlpha) * dt;
    int toiIndexA = bA.islandIndex, toiIndexB = bB.islandIndex;
    // ---- b2Island::SolveTOI ----
    {
      int nb = (int)islandBodies.size();
      std::vector<Position> positions(nb);
      std::vector<Velocity> velocities(nb);
      for (int i = 0; i < nb; ++i) {
        Body& b = bodies[islandBodies[i]];
        positions[i].c = b.sweep.c;
        positions[i].a = b.sweep.a;
        velocities[i].v = b.v;
        velocities[i].w = b.w;
      }
      ContactSolver solver(this, &islandContacts, &positions, &velocities, false, 1.0f);
      for (int i = 0; i < 20; ++i) {
        bool contactsOkay = solver.solvePositionConstraints(true, toiIndexA, toiIndexB);
        if (contactsOkay) break;
      }
      bodies[islandBodies[toiIndexA]].sweep.c0 = positions[toiIndexA].c;
      bodies[islandBodies[toiIndexA]].sweep.a0 = positions[toiIndexA].a;
      bodies[islandBodies[toiIndexB]].sweep.c0 = positions[toiIndexB].c;
      bodies[islandBodies[toiIndexB]].sweep.a0 = positions[toiIndexB].a;
      solver.initializeVelocityConstraints();
      for (int i = 0; i < velIters; ++i) solver.solveVelocityConstraints();
      float h = subDt;
      for (int i = 0; i < nb; ++i) {
        V2 c = positions[i].c;
        float a = positions[i].a;
        V2 v = velocities[i].v;
        float w = velocities[i].w;
        V2 translation = h * v;
        if (dot(translation, translation) > kMaxTranslationSquared) {
          float ratio = kMaxTranslation / length(translation);
          v *= ratio;
        }
        float rotation = h * w;
        if (rotation * rotation > kMaxRotationSquared) {
          float ratio = kMaxRotation / fabs2(rotation);
          w *= ratio;
        }
        c += h * v;
        a += h * w;
        Body& body = bodies[islandBodies[i]];
        body.sweep.c = c;
        body.sweep.a = a;
        body.v = v;
        body.w = w;
        body.synchronizeTransform();
      }
    }
    for (size_t i = 0; i < islandBodies.size(); ++i) {
      int bi = islandBodies[i];
      Body& body = bodies[bi];
      body.islandFlag = false;
      if (body.type != BODY_DYNAMIC) continue;
      synchronizeFixtures(bi);
      for (Contact* c : contacts) {
        int cA = fixtures[c->fA].body, cB = fixtures[c->fB].body;
        if (cA != bi && cB != bi) continue;
        c->toiFlag = false;
        c->islandFlag = false;
      }
    }
    findNewContacts();
  }
}

void World::step(float dt, int velocityIterations, int positionIterations) {
  // e_newFixture: contacts for fixtures created since the last step
  if (newFixture) {
    findNewContacts();
    newFixture = false;
  }
  collide();
  solve(dt, velocityIterations, positionIterations);
  solveTOI(dt, velocityIterations);
  clearForces();
}

}  // namespace b2mini
