// oracle/hockey_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's gym-facing hot path, hockey/hockey_env.py, on top of the
// b2mini engine restatement (b2mini.h).  Each function cites the reference lines it follows.
// It is the CHECKER for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may load it.  Scalar, one env at a time, the reference's own
// float32/float64 mix (NumPy-2 / NEP-50 scalar rules -- the notebook prints np.float64(...)).
//
// Randomness: the reference draws reset positions from gymnasium's PCG64 (hockey_env.py:157-160,
// 180-181) and BasicOpponent phases from the global numpy RNG (hockey_env.py:785,796); neither
// stream is reproducible across implementations, so BOTH this oracle and the CUDA path draw from
// Philox4x32-10 keyed on (seed, global env id, episode/tick) -- same distributions, same draw order.
#include "b2mini.h"
#include "../include/hockey_b200.h"

#include <cstdio>
#include <thread>

using namespace b2mini;

namespace {

// ---- reference constants (hockey_env.py:17-37) -------------------------------------------------
const int FPS = 50;
const double SCALE = 60.0;
const double VIEWPORT_W = 600, VIEWPORT_H = 480;
const double W = VIEWPORT_W / SCALE, H = VIEWPORT_H / SCALE;
const double CENTER_X = W / 2, CENTER_Y = H / 2;
const double ZONE = W / 20;
const double MAX_ANGLE = M_PI / 3;
const int MAX_TIME_KEEP_PUCK = 15;
const double GOAL_SIZE = 75;
const double RACKETPOLY[7][2] = {{-10, 20}, {+5, 20}, {+5, -20}, {-10, -20}, {-18, -10}, {-21, 0}, {-18, 10}};
const double RACKETFACTOR = 1.2;
const int FORCEMULTIPLIER = 6000;
const int SHOOTFORCEMULTIPLIER = 60;
const int TORQUEMULTIPLIER = 400;
const double MAX_PUCK_SPEED = 25;

// ---- Philox4x32-10 ------------------------------------------------------------------------------
struct U4 {
  uint32_t x, y, z, w;
};
static inline U4 philox(uint64_t seed, uint64_t env, uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)env, c1 = (uint32_t)(env >> 32);
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  U4 r = {c0, c1, c2, c3};
  return r;
}
static inline double u53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
static inline float u_pm1(uint32_t x) { return (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f; }
enum { STREAM_RESET = 0, STREAM_OPP = 1, STREAM_ACT = 2, STREAM_PHASE0 = 3 };

// fixture / body numbering (creation order of hockey_env.py:303-317, 373-406 without the
// non-colliding decorations of :239-301, which have categoryBits = maskBits = 0)
enum { F_WALL0 = 0, F_G1_SENSOR = 6, F_G1_SOLID = 7, F_G2_SENSOR = 8, F_G2_SOLID = 9, F_R1 = 10, F_R2 = 11, F_PUCK = 12 };
enum { B_WALL0 = 0, B_GOAL1 = 6, B_GOAL2 = 7, B_R1 = 8, B_R2 = 9, B_PUCK = 10 };

static int pairId(int fA, int fB) {
  int lo = fA < fB ? fA : fB, hi = fA < fB ? fB : fA;
  auto statIdx = [](int f) -> int {
    if (f <= 5) return f;
    if (f == F_G1_SOLID) return 6;
    if (f == F_G2_SOLID) return 7;
    return -1;
  };
  if (hi == F_R1 && statIdx(lo) >= 0) return statIdx(lo);
  if (hi == F_R2 && lo != F_R1 && statIdx(lo) >= 0) return 8 + statIdx(lo);
  if (lo == F_R1 && hi == F_R2) return 16;
  if (hi == F_PUCK) {
    if (lo <= 5) return 17 + lo;
    if (lo == F_G1_SENSOR) return 23;
    if (lo == F_G2_SENSOR) return 24;
    if (lo == F_R1) return 25;
    if (lo == F_R2) return 26;
  }
  return -1;
}
static void pairFixtures(int pid, int* fA, int* fB) {
  static const int stat[8] = {0, 1, 2, 3, 4, 5, F_G1_SOLID, F_G2_SOLID};
  if (pid < 8) { *fA = stat[pid]; *fB = F_R1; }
  else if (pid < 16) { *fA = stat[pid - 8]; *fB = F_R2; }
  else if (pid == 16) { *fA = F_R1; *fB = F_R2; }
  else if (pid < 23) { *fA = pid - 17; *fB = F_PUCK; }
  else if (pid == 23) { *fA = F_G1_SENSOR; *fB = F_PUCK; }
  else if (pid == 24) { *fA = F_G2_SENSOR; *fB = F_PUCK; }
  else if (pid == 25) { *fA = F_R1; *fB = F_PUCK; }
  else { *fA = F_R2; *fB = F_PUCK; }
}

struct Env;
static void onBeginContact(void* user, const Contact* c);

struct Env {
  World world;
  int mode, keep_mode;
  uint64_t seed, env_id;
  int has1, has2;
  bool done, one_starts;
  int winner, time, max_timesteps;
  double phase[2];
  uint32_t episode, tick;
  double ret[2];

  Body& body(int b) { return world.bodies[b]; }

  // ---- _create_world / _create_goal / _create_player / _create_puck (hockey_env.py:183-343) ----
  static Shape polyFrom(const double (*pts)[2], int n, double sx, double sy, double scale) {
    V2 vs[16];
    for (int i = 0; i < n; ++i) vs[i] = mk((float)(sx * pts[i][0] / scale), (float)(sy * pts[i][1] / scale));
    Shape s;
    std::memset(&s, 0, sizeof(s));
    s.setPolygon(vs, n);
    return s;
  }
  void createWall(double px, double py, const double (*poly)[2], double sx, double sy) {
    int b = world.createBody(BODY_STATIC, mk((float)px, (float)py), 0.0f);
    Shape s = polyFrom(poly, 4, sx, sy, SCALE);
    world.createFixture(b, s, 0.0f, 0.1f, 0.0f, 0x011, 0x0011, false);
  }
  void createGoal(double px, double py) {
    const double poly[4][2] = {{-10, GOAL_SIZE}, {10, GOAL_SIZE}, {10, -GOAL_SIZE}, {-10, -GOAL_SIZE}};
    int b = world.createBody(BODY_STATIC, mk((float)px, (float)py), 0.0f);
    Shape s = polyFrom(poly, 4, 1, 1, SCALE);
    world.createFixture(b, s, 0.0f, 0.1f, 0.0f, 0x0010, 0x001, true);
    world.createFixture(b, s, 0.0f, 0.1f, 0.0f, 0x010, 0x0010, false);
  }
  void createPlayer(double px, double py, bool is_two) {
    int b = world.createBody(BODY_DYNAMIC, mk((float)px, (float)py), 0.0f);
    V2 vs[7];
    for (int i = 0; i < 7; ++i) {
      double x = is_two ? -RACKETPOLY[i][0] / SCALE * RACKETFACTOR : RACKETPOLY[i][0] / SCALE * RACKETFACTOR;
      double y = RACKETPOLY[i][1] / SCALE * RACKETFACTOR;
      vs[i] = mk((float)x, (float)y);
    }
    Shape s;
    std::memset(&s, 0, sizeof(s));
    s.setPolygon(vs, 7);
    world.createFixture(b, s, (float)(200.0 / RACKETFACTOR), 1.0f, 0.0f, 0x0010, 0x011, false);
  }
  void createPuck(double px, double py) {
    int b = world.createBody(BODY_DYNAMIC, mk((float)px, (float)py), 0.0f);
    Shape s;
    std::memset(&s, 0, sizeof(s));
    s.setCircle((float)(13 / SCALE));
    world.createFixture(b, s, 7.0f, 0.1f, 0.95f, 0x001, 0x0010, false);
    body(b).linearDamping = 0.05f;
  }

  const double* forced_draws = nullptr;  // tests: replace the reset draws (values, in draw order)
  int64_t reset_seed = -1;  // >= 0: reset(seed=...) (hockey_env.py:347): the draws depend on this seed alone
  double r_uniform(double lo, double hi, int idx) {
    if (forced_draws) return forced_draws[idx];
    U4 r = reset_seed >= 0 ? philox((uint64_t)reset_seed, 0x5EEDED5EEDull, 0u, (uint32_t)STREAM_RESET | ((uint32_t)(idx >> 1) << 8))
                           : philox(seed, env_id, episode, (uint32_t)STREAM_RESET | ((uint32_t)(idx >> 1) << 8));
    double u = (idx & 1) ? u53(r.z, r.w) : u53(r.x, r.y);
    return lo + (hi - lo) * u;
  }

  // ---- reset (hockey_env.py:345-418) ----
  void reset(int one_starting /* -1 = alternate */) {
    world.clear();
    world.beginContact = onBeginContact;
    world.listenerUser = this;
    done = false;
    winner = 0;
    time = 0;
    // NOTE: the reference does not clear player{1,2}_has_puck on reset (hockey_env.py:345-418)
    if (mode == HK_MODE_NORMAL) {
      max_timesteps = 250;
      if (one_starting >= 0)
        one_starts = one_starting != 0;
      else
        one_starts = !one_starts;
    } else {
      max_timesteps = 80;
    }
    // walls (hockey_env.py:307-317)
    const double poly[4][2] = {{-250, 10}, {-250, -10}, {250, -10}, {250, 10}};
    createWall(W / 2, H - .5, poly, 1, 1);
    createWall(W / 2, .5, poly, 1, 1);
    const double cp[4][2] = {{-10, (H - 1) / 2 * SCALE - GOAL_SIZE}, {10, (H - 1) / 2 * SCALE - GOAL_SIZE - 7}, {10, -5}, {-10, -5}};
    createWall(W / 2 - 245 / SCALE, H - .5, cp, 1, -1);
    createWall(W / 2 - 245 / SCALE, .5, cp, 1, 1);
    createWall(W / 2 + 245 / SCALE, H - .5, cp, -1, -1);
    createWall(W / 2 + 245 / SCALE, 0.5, cp, -1, 1);
    // goals (hockey_env.py:373-375)
    createGoal(W / 2 - 245 / SCALE - 10 / SCALE, H / 2);
    createGoal(W / 2 + 245 / SCALE + 10 / SCALE, H / 2);
    // players (hockey_env.py:379-396)
    createPlayer(W / 5, H / 2, false);
    int draw = 0;
    if (mode != HK_MODE_NORMAL) {
      double dx = r_uniform(-W / 3, W / 6, draw++);
      double dy = r_uniform(-H / 4, H / 4, draw++);
      createPlayer(4 * W / 5 + dx, H / 2 + dy, true);
    } else {
      createPlayer(4 * W / 5, H / 2, true);
    }
    // puck (hockey_env.py:397-411)
    if (mode == HK_MODE_NORMAL || mode == HK_MODE_TRAIN_SHOOTING) {
      double dx = r_uniform(H / 8, H / 4, draw++);
      double dy = r_uniform(-H / 8, H / 8, draw++);
      if (one_starts || mode == HK_MODE_TRAIN_SHOOTING)
        createPuck(W / 2 - dx, H / 2 + dy);
      else
        createPuck(W / 2 + dx, H / 2 + dy);
    } else {
      double dx = r_uniform(0, W / 3, draw++);
      double dy = r_uniform(-H / 2, H / 2, draw++);
      createPuck(W / 2 + dx, H / 2 + 0.8 * dy);
      double ay = r_uniform(-GOAL_SIZE / SCALE, GOAL_SIZE / SCALE, draw++);
      Body& puck = body(B_PUCK);
      // b2Vec2 arithmetic is float32 (pybox2d): direction = position - (0, H/2 + .6*r)
      V2 direction = puck.xf.p - mk(0.0f, (float)(H / 2 + .6 * ay));
      float len = length(direction);
      direction = mk(direction.x / len, direction.y / len);
      V2 force = -direction;
      force = mk(force.x * (float)SHOOTFORCEMULTIPLIER, force.y * (float)SHOOTFORCEMULTIPLIER);
      force = mk(force.x * puck.mass, force.y * puck.mass);
      float ts = (float)(1.0 / FPS);
      force = mk(force.x / ts, force.y / ts);
      puck.applyForceToCenter(force, true);
    }
    ++episode;
    ret[0] = ret[1] = 0.0;
  }

  // ---- _get_obs / obs_agent_two (hockey_env.py:485-516) ----
  void getObs(float* o) {
    Body &p1 = body(B_R1), &p2 = body(B_R2), &pk = body(B_PUCK);
    const float cx = (float)CENTER_X, cy = (float)CENTER_Y;
    o[0] = p1.xf.p.x - cx; o[1] = p1.xf.p.y - cy; o[2] = p1.sweep.a;
    o[3] = p1.v.x; o[4] = p1.v.y; o[5] = p1.w;
    o[6] = p2.xf.p.x - cx; o[7] = p2.xf.p.y - cy; o[8] = p2.sweep.a;
    o[9] = p2.v.x; o[10] = p2.v.y; o[11] = p2.w;
    o[12] = pk.xf.p.x - cx; o[13] = pk.xf.p.y - cy; o[14] = pk.v.x; o[15] = pk.v.y;
    o[16] = (float)has1; o[17] = (float)has2;
  }
  void getObs2(float* o) {
    Body &p1 = body(B_R1), &p2 = body(B_R2), &pk = body(B_PUCK);
    const float cx = (float)CENTER_X, cy = (float)CENTER_Y;
    o[0] = -(p2.xf.p.x - cx); o[1] = -(p2.xf.p.y - cy); o[2] = p2.sweep.a;
    o[3] = -p2.v.x; o[4] = -p2.v.y; o[5] = p2.w;
    o[6] = -(p1.xf.p.x - cx); o[7] = -(p1.xf.p.y - cy); o[8] = p1.sweep.a;
    o[9] = -p1.v.x; o[10] = -p1.v.y; o[11] = p1.w;
    o[12] = -(pk.xf.p.x - cx); o[13] = -(pk.xf.p.y - cy); o[14] = -pk.v.x; o[15] = -pk.v.y;
    o[16] = (float)has2; o[17] = (float)has1;
  }

  // ---- _get_info / get_info_agent_two (hockey_env.py:542-591); out = winner, closeness, touch, direction
  void getInfo(double* out, bool agent_two) {
    Body &p1 = body(B_R1), &p2 = body(B_R2), &pk = body(B_PUCK);
    double closeness = 0;
    bool cond = agent_two ? ((double)pk.xf.p.x > CENTER_X && (double)pk.v.x >= 0)
                          : ((double)pk.xf.p.x < CENTER_X && (double)pk.v.x <= 0);
    if (cond) {
      V2 d = (agent_two ? p2.xf.p : p1.xf.p) - pk.xf.p;  // b2Vec2 subtraction: float32
      double dist = std::sqrt((double)d.x * (double)d.x + (double)d.y * (double)d.y);
      double max_dist = 250. / SCALE;
      double max_reward = -30.;
      double factor = max_reward / (max_dist * max_timesteps / 2);
      closeness += dist * factor;
    }
    double touch = 0.;
    if ((agent_two ? has2 : has1) == MAX_TIME_KEEP_PUCK) touch = 1.;
    double factor = (agent_two ? -1.0 : 1.0) / (max_timesteps * MAX_PUCK_SPEED);
    double direction = (double)pk.v.x * factor;
    out[0] = agent_two ? -winner : winner;
    out[1] = closeness;
    out[2] = touch;
    out[3] = direction;
  }
  double computeReward() {  // hockey_env.py:518-528
    double r = 0;
    if (done) {
      if (winner == 1) r += 10;
      else if (winner == -1) r -= 10;
    }
    return r;
  }

  // ---- _check_boundaries (hockey_env.py:420-434); force values are float32-exact ----
  void checkBoundaries(double force[2], Body& player, bool is_one) {
    double px = player.xf.p.x, py = player.xf.p.y;
    if ((is_one && px < W / 2 - 210 / SCALE && force[0] < 0) || (!is_one && px > W / 2 + 210 / SCALE && force[0] > 0) ||
        (is_one && px > W / 2 && force[0] > 0) || (!is_one && px < W / 2 && force[0] < 0)) {
      player.v.x = 0;  // SWIG reference proxy: writes through to the body
      force[0] = -(double)player.v.x;
    }
    if ((py > H - 1.2 && force[1] > 0) || (py < 1.2 && force[1] < 0)) {
      player.v.y = 0;
      force[1] = -(double)player.v.y;
    }
  }

  // ---- _apply_translation_action_with_max_speed (hockey_env.py:436-470) ----
  void applyTranslation(Body& player, const float action[2], double max_speed, bool is_one) {
    const double timeStep = 1.0 / FPS;
    double vel[2] = {(double)player.v.x, (double)player.v.y};
    double speed = std::sqrt(vel[0] * vel[0] + vel[1] * vel[1]);
    float force[2];
    if (is_one) {
      force[0] = action[0] * (float)FORCEMULTIPLIER;
      force[1] = action[1] * (float)FORCEMULTIPLIER;
    } else {
      force[0] = (-action[0]) * (float)FORCEMULTIPLIER;
      force[1] = (-action[1]) * (float)FORCEMULTIPLIER;
    }
    double px = player.xf.p.x, vx = player.v.x, mass = player.mass;
    if ((is_one && px > CENTER_X - ZONE) || (!is_one && px < CENTER_X + ZONE)) {
      force[0] = 0;
      if (is_one) {
        if (vx > 0) force[0] = (float)(-2 * vx * mass / timeStep);
        force[0] += (float)(-1 * (px - CENTER_X) * vx * mass / timeStep);
      } else {
        if (vx < 0) force[0] = (float)(-2 * vx * mass / timeStep);
        force[0] += (float)(1 * (px - CENTER_X) * vx * mass / timeStep);
      }
      player.linearDamping = 20.0f;
      double f[2] = {force[0], force[1]};
      checkBoundaries(f, player, is_one);
      player.applyForceToCenter(mk((float)f[0], (float)f[1]), true);
      return;
    }
    if (speed < max_speed) {
      player.linearDamping = 5.0f;
      double f[2] = {force[0], force[1]};
      checkBoundaries(f, player, is_one);
      player.applyForceToCenter(mk((float)f[0], (float)f[1]), true);
    } else {
      player.linearDamping = 20.0f;
      // deltaVelocity = self.timeStep * force / player.mass : float32 array arithmetic
      float ts = (float)timeStep, m32 = player.mass;
      float dv0 = (ts * force[0]) / m32, dv1 = (ts * force[1]) / m32;
      double n0 = vel[0] + (double)dv0, n1 = vel[1] + (double)dv1;
      if (std::sqrt(n0 * n0 + n1 * n1) < speed) {
        double f[2] = {force[0], force[1]};
        checkBoundaries(f, player, is_one);
        player.applyForceToCenter(mk((float)f[0], (float)f[1]), true);
      }
    }
  }

  // ---- _apply_rotation_action_with_max_speed (hockey_env.py:472-483) ----
  void applyRotation(Body& player, float action) {
    const double timeStep = 1.0 / FPS;
    double angle = player.sweep.a;
    double torque = (double)(action * (float)TORQUEMULTIPLIER);
    if (std::fabs(angle) > MAX_ANGLE) {
      torque = 0;
      if (angle * (double)player.w > 0) torque = -0.1 * (double)player.w * (double)player.mass / timeStep;
      torque += -0.1 * angle * (double)player.mass / timeStep;
      player.angularDamping = 10.0f;
    } else {
      player.angularDamping = 2.0f;
    }
    player.applyTorque((float)torque, true);
  }

  // ---- _keep_puck / _shoot (hockey_env.py:618-633) ----
  void keepPuck(Body& player) {
    world.setTransform(B_PUCK, player.xf.p, body(B_PUCK).sweep.a);
    body(B_PUCK).setLinearVelocity(player.v);
  }
  void shoot(Body& player, bool is_one) {
    Body& puck = body(B_PUCK);
    double a = player.sweep.a;
    double c, s;
    if (g_trig_mode == 1) { c = std::cos(a); s = std::sin(a); } else { sincos_poly(a, &s, &c); }
    double sgn = is_one ? 1.0 : -1.0;
    V2 f = mk((float)(c * sgn), (float)(s * sgn));
    f = mk(f.x * puck.mass, f.y * puck.mass);
    float ts = (float)(1.0 / FPS);
    f = mk(f.x / ts, f.y / ts);
    f = mk(f.x * (float)SHOOTFORCEMULTIPLIER, f.y * (float)SHOOTFORCEMULTIPLIER);
    puck.applyForceToCenter(f, true);
  }

  // ---- step (hockey_env.py:658-695); action already clipped float32 ----
  void step(const float action[8], float* obs, double* reward, double* info /*4*/) {
    Body &p1 = body(B_R1), &p2 = body(B_R2), &puck = body(B_PUCK);
    applyTranslation(p1, action + 0, 10, true);
    applyRotation(p1, action[2]);
    applyTranslation(p2, action + 4, 10, false);
    applyRotation(p2, action[6]);
    {  // _limit_puck_speed (hockey_env.py:610-616)
      double vx = puck.v.x, vy = puck.v.y;
      double puck_speed = std::sqrt(vx * vx + vy * vy);
      puck.linearDamping = puck_speed > MAX_PUCK_SPEED ? 10.0f : 0.05f;
    }
    if (keep_mode) {
      if (has1 > 1) {
        keepPuck(p1);
        has1 -= 1;
        if (has1 == 1 || action[3] > 0.5f) {
          shoot(p1, true);
          has1 = 0;
        }
      }
      if (has2 > 1) {
        keepPuck(p2);
        has2 -= 1;
        if (has2 == 1 || action[7] > 0.5f) {
          shoot(p2, false);
          has2 = 0;
        }
      }
    }
    world.step((float)(1.0 / FPS), 6 * 30, 2 * 30);
    if (obs) getObs(obs);
    if (time >= max_timesteps) done = true;
    getInfo(info, false);
    *reward = computeReward() + info[1];
    time += 1;
    ++tick;
  }

  // ---- BasicOpponent.act (hockey_env.py:787-833) on a float32 observation ----
  void basicAct(const float* obs, bool weak, double* phase_io, double u_inc, float out[4]) {
    double p1[3] = {obs[0], obs[1], obs[2]};
    double v1[3] = {obs[3], obs[4], obs[5]};
    double puck[2] = {obs[12], obs[13]};
    double puckv[2] = {obs[14], obs[15]};
    double target_pos[2];
    *phase_io += u_inc;
    const double time_to_break = 0.1;
    double kp = weak ? 0.5 : 10;
    const double kd = 0.5;
    if (puckv[0] < 30.0 / SCALE) {
      double dx = p1[0] - puck[0], dy = p1[1] - puck[1];
      double dist = std::sqrt(dx * dx + dy * dy);
      if (p1[0] < puck[0] && std::fabs(p1[1] - puck[1]) < 30.0 / SCALE) {
        target_pos[0] = puck[0] + 0.2;
        target_pos[1] = puck[1] + puckv[1] * dist * 0.1;
      } else {
        target_pos[0] = -210 / SCALE;
        target_pos[1] = puck[1];
      }
    } else {
      target_pos[0] = -210 / SCALE;
      target_pos[1] = 0;
    }
    double sp;
    if (g_trig_mode == 1) sp = std::sin(*phase_io); else { double cdummy; sincos_poly(*phase_io, &sp, &cdummy); }
    double target_angle = MAX_ANGLE * sp;
    double shoot = 0.0;
    if (keep_mode && obs[16] > 0 && obs[16] < 7) shoot = 1.0;
    double target[3] = {target_pos[0], target_pos[1], target_angle};
    double kps[3] = {kp, kp / 5, kp / 2};
    double ttb[3] = {time_to_break, time_to_break, time_to_break * 10};
    for (int i = 0; i < 3; ++i) {
      double error = target[i] - p1[i];
      double need_break = std::fabs(error / (v1[i] + 0.01)) < ttb[i] ? 1.0 : 0.0;
      double a = error * kps[i] - v1[i] * need_break * kd;
      a = a < -1 ? -1 : (a > 1 ? 1 : a);  // np.clip
      out[i] = (float)a;                  // step(): np.clip(action,-1,1).astype(np.float32)
    }
    out[3] = (float)shoot;
  }
};

// ---- ContactDetector.BeginContact (hockey_env.py:50-73) ----
static void onBeginContact(void* user, const Contact* c) {
  Env* env = (Env*)user;
  int bA = env->world.fixtures[c->fA].body, bB = env->world.fixtures[c->fB].body;
  if (bA == B_GOAL2 || bB == B_GOAL2) {
    if (bA == B_PUCK || bB == B_PUCK) {
      env->done = true;
      env->winner = 1;
    }
  }
  if (bA == B_GOAL1 || bB == B_GOAL1) {
    if (bA == B_PUCK || bB == B_PUCK) {
      env->done = true;
      env->winner = -1;
    }
  }
  if ((bA == B_R1 || bB == B_R1) && (bA == B_PUCK || bB == B_PUCK)) {
    if (env->keep_mode && (double)env->body(B_PUCK).v.x < 0.1) {
      if (env->has1 == 0) env->has1 = MAX_TIME_KEEP_PUCK;
    }
  }
  if ((bA == B_R2 || bB == B_R2) && (bA == B_PUCK || bB == B_PUCK)) {
    if (env->keep_mode && (double)env->body(B_PUCK).v.x > -0.1) {
      if (env->has2 == 0) env->has2 = MAX_TIME_KEEP_PUCK;
    }
  }
}

struct Batch {
  std::vector<Env*> envs;
  int n_threads;
  double stats[HK_STATS_DIM];
};

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static void packState(Env* e, uint32_t* r) {
  std::memset(r, 0, sizeof(uint32_t) * HK_STATE_WORDS);
  for (int k = 0; k < 2; ++k) {
    Body& b = e->body(k == 0 ? B_R1 : B_R2);
    uint32_t* p = r + (k == 0 ? HK_S_R1 : HK_S_R2);
    p[0] = f2u(b.xf.p.x); p[1] = f2u(b.xf.p.y); p[2] = f2u(b.sweep.c.x); p[3] = f2u(b.sweep.c.y);
    p[4] = f2u(b.sweep.a); p[5] = f2u(b.v.x); p[6] = f2u(b.v.y); p[7] = f2u(b.w);
  }
  Body& pk = e->body(B_PUCK);
  uint32_t* p = r + HK_S_PUCK;
  p[0] = f2u(pk.sweep.c.x); p[1] = f2u(pk.sweep.c.y); p[2] = f2u(pk.sweep.a);
  p[3] = f2u(pk.v.x); p[4] = f2u(pk.v.y); p[5] = f2u(pk.w);
  r[HK_S_SLEEP + 0] = f2u(e->body(B_R1).sleepTime);
  r[HK_S_SLEEP + 1] = f2u(e->body(B_R2).sleepTime);
  r[HK_S_SLEEP + 2] = f2u(pk.sleepTime);
  uint32_t flags = 0;
  if (e->body(B_R1).awake) flags |= 1;
  if (e->body(B_R2).awake) flags |= 2;
  if (pk.awake) flags |= 4;
  if (e->done) flags |= 8;
  if (e->one_starts) flags |= 16;
  flags |= (uint32_t)(e->winner + 1) << 5;
  r[HK_S_FLAGS] = flags;
  r[HK_S_TIME] = (uint32_t)e->time;
  r[HK_S_HAS1] = (uint32_t)e->has1;
  r[HK_S_HAS2] = (uint32_t)e->has2;
  r[HK_S_PFORCE] = f2u(pk.force.x);
  r[HK_S_PFORCE + 1] = f2u(pk.force.y);
  const int fx[3] = {F_R1, F_R2, F_PUCK};
  for (int k = 0; k < 3; ++k) {
    const AABB& a = e->world.fixtures[fx[k]].fatAABB;
    r[HK_S_FAT + 4 * k + 0] = f2u(a.lo.x); r[HK_S_FAT + 4 * k + 1] = f2u(a.lo.y);
    r[HK_S_FAT + 4 * k + 2] = f2u(a.hi.x); r[HK_S_FAT + 4 * k + 3] = f2u(a.hi.y);
  }
  uint32_t moved = 0;
  for (int f : e->world.moveBuffer) {
    if (f == F_R1) moved |= 1;
    if (f == F_R2) moved |= 2;
    if (f == F_PUCK) moved |= 4;
  }
  if (e->world.newFixture) moved |= 8;
  r[HK_S_MOVED] = moved;
  std::memcpy(r + HK_S_PHASE, e->phase, 16);
  r[HK_S_EPISODE] = e->episode;
  r[HK_S_TICK] = e->tick;
  std::memcpy(r + HK_S_RET, e->ret, 16);
  r[HK_S_PUCK_C0] = f2u(pk.sweep.c0.x);
  r[HK_S_PUCK_C0 + 1] = f2u(pk.sweep.c0.y);
  for (size_t i = 0; i < e->world.contacts.size(); ++i) {
    Contact* c = e->world.contacts[i];
    int pid = pairId(c->fA, c->fB);
    if (pid < 0) continue;
    uint32_t* q = r + HK_S_CONTACT + HK_CONTACT_WORDS * pid;
    q[0] = 1u | (c->touching ? 2u : 0u) | ((uint32_t)i << 8);
    q[1] = (uint32_t)c->manifold.pointCount;
    for (int j = 0; j < c->manifold.pointCount; ++j) {
      q[2 + j] = c->manifold.points[j].key;
      q[4 + 2 * j] = f2u(c->manifold.points[j].normalImpulse);
      q[5 + 2 * j] = f2u(c->manifold.points[j].tangentImpulse);
    }
  }
}

static void unpackState(Env* e, const uint32_t* r) {
  // bodies must already exist (reset() built the scene); overwrite dynamic state
  for (int k = 0; k < 2; ++k) {
    Body& b = e->body(k == 0 ? B_R1 : B_R2);
    const uint32_t* p = r + (k == 0 ? HK_S_R1 : HK_S_R2);
    b.xf.p = mk(u2f(p[0]), u2f(p[1]));
    b.sweep.c = b.sweep.c0 = mk(u2f(p[2]), u2f(p[3]));
    b.sweep.a = b.sweep.a0 = u2f(p[4]);
    b.xf.q.set(b.sweep.a);
    b.v = mk(u2f(p[5]), u2f(p[6]));
    b.w = u2f(p[7]);
    b.force = mk(0, 0);
    b.torque = 0;
  }
  Body& pk = e->body(B_PUCK);
  const uint32_t* p = r + HK_S_PUCK;
  pk.sweep.c = mk(u2f(p[0]), u2f(p[1]));
  pk.sweep.c0 = mk(u2f(r[HK_S_PUCK_C0]), u2f(r[HK_S_PUCK_C0 + 1]));
  pk.xf.p = pk.sweep.c;
  pk.sweep.a = pk.sweep.a0 = u2f(p[2]);
  pk.xf.q.set(pk.sweep.a);
  pk.v = mk(u2f(p[3]), u2f(p[4]));
  pk.w = u2f(p[5]);
  e->body(B_R1).sleepTime = u2f(r[HK_S_SLEEP + 0]);
  e->body(B_R2).sleepTime = u2f(r[HK_S_SLEEP + 1]);
  pk.sleepTime = u2f(r[HK_S_SLEEP + 2]);
  uint32_t flags = r[HK_S_FLAGS];
  e->body(B_R1).awake = flags & 1;
  e->body(B_R2).awake = flags & 2;
  pk.awake = flags & 4;
  e->done = flags & 8;
  e->one_starts = flags & 16;
  e->winner = (int)((flags >> 5) & 3) - 1;
  e->time = (int)r[HK_S_TIME];
  e->has1 = (int)r[HK_S_HAS1];
  e->has2 = (int)r[HK_S_HAS2];
  pk.force = mk(u2f(r[HK_S_PFORCE]), u2f(r[HK_S_PFORCE + 1]));
  pk.torque = 0;
  const int fx[3] = {F_R1, F_R2, F_PUCK};
  for (int k = 0; k < 3; ++k) {
    AABB& a = e->world.fixtures[fx[k]].fatAABB;
    a.lo = mk(u2f(r[HK_S_FAT + 4 * k + 0]), u2f(r[HK_S_FAT + 4 * k + 1]));
    a.hi = mk(u2f(r[HK_S_FAT + 4 * k + 2]), u2f(r[HK_S_FAT + 4 * k + 3]));
  }
  e->world.moveBuffer.clear();
  uint32_t moved = r[HK_S_MOVED];
  if (moved & 8) {
    for (int f = 0; f <= F_PUCK; ++f) e->world.moveBuffer.push_back(f);
  } else {
    if (moved & 1) e->world.moveBuffer.push_back(F_R1);
    if (moved & 2) e->world.moveBuffer.push_back(F_R2);
    if (moved & 4) e->world.moveBuffer.push_back(F_PUCK);
  }
  e->world.newFixture = (moved & 8) != 0;
  std::memcpy(e->phase, r + HK_S_PHASE, 16);
  e->episode = r[HK_S_EPISODE];
  e->tick = r[HK_S_TICK];
  std::memcpy(e->ret, r + HK_S_RET, 16);
  // contacts: rebuild the world contact list in recorded order
  for (Contact* c : e->world.contacts) delete c;
  e->world.contacts.clear();
  std::vector<std::pair<int, int>> order;
  for (int pid = 0; pid < HK_N_PAIRS; ++pid) {
    const uint32_t* q = r + HK_S_CONTACT + HK_CONTACT_WORDS * pid;
    if (q[0] & 1u) order.push_back(std::make_pair((int)((q[0] >> 8) & 255), pid));
  }
  std::sort(order.begin(), order.end());
  for (auto& op : order) {
    int pid = op.second;
    const uint32_t* q = r + HK_S_CONTACT + HK_CONTACT_WORDS * pid;
    Contact* c = new Contact();
    pairFixtures(pid, &c->fA, &c->fB);
    std::memset(&c->manifold, 0, sizeof(c->manifold));
    c->touching = (q[0] & 2u) != 0;
    c->enabled = true;
    c->islandFlag = false;
    c->toiFlag = false;
    c->toiCount = 0;
    c->toi = 1.0f;
    const Fixture& fa = e->world.fixtures[c->fA];
    const Fixture& fb = e->world.fixtures[c->fB];
    c->friction = sqrtf(fa.friction * fb.friction);
    c->restitution = fa.restitution > fb.restitution ? fa.restitution : fb.restitution;
    // geometry is re-derived from the current poses (only read again if Collide skips this
    // contact because both bodies sleep); ids and impulses come from the record.
    if (!fa.isSensor && !fb.isSensor) {
      Manifold g;
      std::memset(&g, 0, sizeof(g));
      const Xf& xa = e->world.bodies[fa.body].xf;
      const Xf& xb = e->world.bodies[fb.body].xf;
      if (fb.shape.type == SHAPE_CIRCLE) collidePolygonAndCircle(&g, &fa.shape, xa, &fb.shape, xb);
      else collidePolygons(&g, &fa.shape, xa, &fb.shape, xb);
      if (g.pointCount == (int)q[1]) c->manifold = g;
    }
    c->manifold.pointCount = (int)q[1];
    for (int j = 0; j < c->manifold.pointCount; ++j) {
      c->manifold.points[j].key = q[2 + j];
      c->manifold.points[j].normalImpulse = u2f(q[4 + 2 * j]);
      c->manifold.points[j].tangentImpulse = u2f(q[5 + 2 * j]);
    }
    e->world.contacts.push_back(c);
  }
}

static void actionsFor(Env* e, const float* action, int stride, int pol1, int pol2, float a[8]) {
  float obs[18];
  int pol[2] = {pol1, pol2};
  U4 ro = {0, 0, 0, 0};
  bool need_opp = (pol1 == HK_POLICY_BASIC_WEAK || pol1 == HK_POLICY_BASIC_STRONG || pol2 == HK_POLICY_BASIC_WEAK ||
                   pol2 == HK_POLICY_BASIC_STRONG);
  if (need_opp) ro = philox(e->seed, e->env_id, e->tick, STREAM_OPP);
  for (int k = 0; k < 2; ++k) {
    float* out = a + 4 * k;
    switch (pol[k]) {
      case HK_POLICY_EXTERNAL:
        for (int i = 0; i < 4; ++i) {
          float x = action[k * 4 + i];
          out[i] = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x);  // np.clip(action,-1,1).astype(float32)
        }
        break;
      case HK_POLICY_BASIC_WEAK:
      case HK_POLICY_BASIC_STRONG: {
        if (k == 0) e->getObs(obs); else e->getObs2(obs);
        double u = k == 0 ? u53(ro.x, ro.y) : u53(ro.z, ro.w);
        e->basicAct(obs, pol[k] == HK_POLICY_BASIC_WEAK, &e->phase[k], 0.0 + (0.2 - 0.0) * u, out);
      } break;
      case HK_POLICY_RANDOM: {
        U4 r = philox(e->seed, e->env_id, e->tick, (uint32_t)STREAM_ACT | ((uint32_t)k << 8));
        out[0] = u_pm1(r.x); out[1] = u_pm1(r.y); out[2] = u_pm1(r.z); out[3] = u_pm1(r.w);
      } break;
      default:
        out[0] = out[1] = out[2] = out[3] = 0.0f;
    }
  }
  (void)stride;
}

}  // namespace

extern "C" {

void* hko_create(int64_t n, int mode, int keep_mode, uint64_t seed, int64_t env_id_offset, int n_threads) {
  Batch* b = new Batch();
  b->n_threads = n_threads > 0 ? n_threads : 1;
  std::memset(b->stats, 0, sizeof(b->stats));
  for (int64_t i = 0; i < n; ++i) {
    Env* e = new Env();
    e->mode = mode;
    e->keep_mode = keep_mode;
    e->seed = seed;
    e->env_id = (uint64_t)(env_id_offset + i);
    e->has1 = e->has2 = 0;
    e->one_starts = true;
    e->episode = 0;
    e->tick = 0;
    U4 r = philox(seed, e->env_id, 0, STREAM_PHASE0);
    e->phase[0] = 0.0 + (M_PI - 0.0) * u53(r.x, r.y);  // np.random.uniform(0, np.pi) (hockey_env.py:785)
    e->phase[1] = 0.0 + (M_PI - 0.0) * u53(r.z, r.w);
    e->reset(1);
    b->envs.push_back(e);
  }
  return b;
}
void hko_destroy(void* h) {
  Batch* b = (Batch*)h;
  for (Env* e : b->envs) delete e;
  delete b;
}
void hko_set_modes(int trig_mode, int static_drift) {
  g_trig_mode = trig_mode;
  g_static_drift = static_drift;
}
void hko_reset(void* h, const uint8_t* mask, const int8_t* one_starting, float* obs) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) {
    if (mask && !mask[i]) continue;
    b->envs[i]->reset(one_starting ? (int)one_starting[i] : -1);
    if (obs) b->envs[i]->getObs(obs + 18 * i);
  }
}
void hko_reset_seeded(void* h, const uint8_t* mask, const int8_t* one_starting, const int64_t* seeds, float* obs) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) {
    if (mask && !mask[i]) continue;
    b->envs[i]->reset_seed = seeds ? seeds[i] : -1;
    b->envs[i]->reset(one_starting ? (int)one_starting[i] : -1);
    b->envs[i]->reset_seed = -1;
    if (obs) b->envs[i]->getObs(obs + 18 * i);
  }
}
// reset env `index` with the given r_uniform() return values instead of Philox draws (golden tests)
void hko_reset_with_draws(void* h, int64_t index, int one_starting, const double* draws) {
  Batch* b = (Batch*)h;
  Env* e = b->envs[index];
  e->forced_draws = draws;
  e->reset(one_starting);
  e->forced_draws = nullptr;
}
void hko_get_obs(void* h, float* obs, float* obs2) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) {
    if (obs) b->envs[i]->getObs(obs + 18 * i);
    if (obs2) b->envs[i]->getObs2(obs2 + 18 * i);
  }
}

static void stepRange(Batch* b, size_t lo, size_t hi, const float* action, int stride, int pol1, int pol2, int flags,
                      float* obs, float* obs2, double* reward, double* reward2, uint8_t* done, double* info,
                      double* info2, float* final_obs, double* stats) {
  for (size_t i = lo; i < hi; ++i) {
    Env* e = b->envs[i];
    float a[8];
    actionsFor(e, action ? action + (size_t)stride * i : nullptr, stride, pol1, pol2, a);
    double r, inf[4], inf2[4];
    int had1 = e->has1, had2 = e->has2;
    e->step(a, nullptr, &r, inf);
    e->getInfo(inf2, true);
    double r2 = -e->computeReward() + inf2[1];
    e->ret[0] += r;
    e->ret[1] += r2;
    stats[4] += 1;
    if (e->has1 == MAX_TIME_KEEP_PUCK && had1 != MAX_TIME_KEEP_PUCK) stats[9] += 1;
    if (e->has2 == MAX_TIME_KEEP_PUCK && had2 != MAX_TIME_KEEP_PUCK) stats[10] += 1;
    if (reward) reward[i] = r;
    if (reward2) reward2[i] = r2;
    if (done) done[i] = e->done ? 1 : 0;
    if (info) std::memcpy(info + 4 * i, inf, sizeof(inf));
    if (info2) std::memcpy(info2 + 4 * i, inf2, sizeof(inf2));
    if (final_obs) e->getObs(final_obs + 18 * i);
    if (e->done && (flags & HK_STEP_AUTORESET)) {
      stats[0] += 1;
      if (e->winner == 1) stats[1] += 1; else if (e->winner == -1) stats[2] += 1; else stats[3] += 1;
      stats[5] += e->ret[0];
      stats[6] += e->ret[1];
      stats[7] += e->ret[0] * e->ret[0];
      stats[8] += e->time;
      e->reset(-1);
    }
    if (obs) e->getObs(obs + 18 * i);
    if (obs2) e->getObs2(obs2 + 18 * i);
  }
}

void hko_step(void* h, const float* action, int stride, int pol1, int pol2, int flags, float* obs, float* obs2,
              double* reward, double* reward2, uint8_t* done, double* info, double* info2, float* final_obs) {
  Batch* b = (Batch*)h;
  size_t n = b->envs.size();
  int nt = b->n_threads;
  if (nt <= 1 || n < (size_t)(2 * nt)) {
    stepRange(b, 0, n, action, stride, pol1, pol2, flags, obs, obs2, reward, reward2, done, info, info2, final_obs,
              b->stats);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::vector<double>> st(nt, std::vector<double>(HK_STATS_DIM, 0.0));
  for (int t = 0; t < nt; ++t) {
    size_t lo = n * t / nt, hi = n * (t + 1) / nt;
    th.emplace_back([=, &st]() {
      stepRange(b, lo, hi, action, stride, pol1, pol2, flags, obs, obs2, reward, reward2, done, info, info2, final_obs,
                st[t].data());
    });
  }
  for (auto& t : th) t.join();
  for (int t = 0; t < nt; ++t)
    for (int k = 0; k < HK_STATS_DIM; ++k) b->stats[k] += st[t][k];
}

// k fused ticks with in-oracle policies and autoreset (the CPU twin of hk_rollout)
void hko_rollout(void* h, int k_steps, int pol1, int pol2) {
  Batch* b = (Batch*)h;
  size_t n = b->envs.size();
  int nt = b->n_threads;
  auto work = [=](size_t lo, size_t hi, double* stats) {
    for (int s = 0; s < k_steps; ++s)
      stepRange(b, lo, hi, nullptr, 0, pol1, pol2, HK_STEP_AUTORESET, nullptr, nullptr, nullptr, nullptr, nullptr,
                nullptr, nullptr, nullptr, stats);
  };
  if (nt <= 1 || n < (size_t)(2 * nt)) {
    work(0, n, b->stats);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::vector<double>> st(nt, std::vector<double>(HK_STATS_DIM, 0.0));
  for (int t = 0; t < nt; ++t) {
    size_t lo = n * t / nt, hi = n * (t + 1) / nt;
    th.emplace_back([=, &st]() { work(lo, hi, st[t].data()); });
  }
  for (auto& t : th) t.join();
  for (int t = 0; t < nt; ++t)
    for (int k = 0; k < HK_STATS_DIM; ++k) b->stats[k] += st[t][k];
}

void hko_get_state(void* h, uint32_t* rec) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) packState(b->envs[i], rec + (size_t)HK_STATE_WORDS * i);
}
void hko_set_state(void* h, const uint32_t* rec) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) unpackState(b->envs[i], rec + (size_t)HK_STATE_WORDS * i);
}
// HockeyEnv.set_state (hockey_env.py:594-608): 18 visible values (float64 in the reference)
void hko_set_obs_state(void* h, const double* st) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) {
    Env* e = b->envs[i];
    const double* s = st + 18 * i;
    const int bi[2] = {B_R1, B_R2};
    for (int k = 0; k < 2; ++k) {
      const double* q = s + 6 * k;
      e->world.setTransform(bi[k], mk((float)(q[0] + CENTER_X), (float)(q[1] + CENTER_Y)), e->body(bi[k]).sweep.a);
      e->world.setTransform(bi[k], e->body(bi[k]).xf.p, (float)q[2]);
      e->body(bi[k]).setLinearVelocity(mk((float)q[3], (float)q[4]));
      if ((float)q[5] * (float)q[5] > 0.0f) e->body(bi[k]).setAwake(true);
      e->body(bi[k]).w = (float)q[5];
    }
    e->world.setTransform(B_PUCK, mk((float)(s[12] + CENTER_X), (float)(s[13] + CENTER_Y)), e->body(B_PUCK).sweep.a);
    e->body(B_PUCK).setLinearVelocity(mk((float)s[14], (float)s[15]));
    if (e->keep_mode) {
      e->has1 = (int)s[16];
      e->has2 = (int)s[17];
    }
  }
}
void hko_get_stats(void* h, double* out) {
  Batch* b = (Batch*)h;
  std::memcpy(out, b->stats, sizeof(b->stats));
  long long toi = 0;
  for (Env* e : b->envs) toi += e->world.nToiEvents;
  out[12] = (double)toi;
}
void hko_clear_stats(void* h) {
  Batch* b = (Batch*)h;
  std::memset(b->stats, 0, sizeof(b->stats));
  for (Env* e : b->envs) e->world.nToiEvents = e->world.nToiCalls = 0;
}
// scene constants for cross-checking the CUDA library's own tables:
// out[0..] = racket mass, invMass, I, invI, localCenter.x (p1), localCenter.x (p2), puck mass, invMass, I, invI, radius
void hko_scene_constants(void* h, float* out) {
  Batch* b = (Batch*)h;
  Env* e = b->envs[0];
  Body &r1 = e->body(B_R1), &r2 = e->body(B_R2), &pk = e->body(B_PUCK);
  out[0] = r1.mass; out[1] = r1.invMass; out[2] = r1.I; out[3] = r1.invI;
  out[4] = r1.sweep.localCenter.x; out[5] = r1.sweep.localCenter.y;
  out[6] = r2.sweep.localCenter.x; out[7] = r2.sweep.localCenter.y;
  out[8] = pk.mass; out[9] = pk.invMass; out[10] = pk.I; out[11] = pk.invI;
  out[12] = e->world.fixtures[F_PUCK].shape.radius;
  out[13] = r2.mass; out[14] = r2.I; out[15] = r2.invI;
}
// polygon tables: for fixture f (0..11) writes count, then count vertices (x,y), count normals, centroid, body position
int hko_scene_polygon(void* h, int f, float* out) {
  Batch* b = (Batch*)h;
  Env* e = b->envs[0];
  const Fixture& fx = e->world.fixtures[f];
  int n = fx.shape.count, k = 0;
  for (int i = 0; i < n; ++i) { out[k++] = fx.shape.v[i].x; out[k++] = fx.shape.v[i].y; }
  for (int i = 0; i < n; ++i) { out[k++] = fx.shape.n[i].x; out[k++] = fx.shape.n[i].y; }
  out[k++] = fx.shape.centroid.x; out[k++] = fx.shape.centroid.y;
  out[k++] = e->world.bodies[fx.body].xf.p.x; out[k++] = e->world.bodies[fx.body].xf.p.y;
  out[k++] = fx.fatAABB.lo.x; out[k++] = fx.fatAABB.lo.y; out[k++] = fx.fatAABB.hi.x; out[k++] = fx.fatAABB.hi.y;
  return n;
}
// float sin/cos of the trig mode in use (for tests/test_oracle_trig.py)
void hko_sincosf(const float* x, int64_t n, float* s, float* c) {
  for (int64_t i = 0; i < n; ++i) sincosf_b2(x[i], s + i, c + i);
}
// sampled check of the polynomial sin/cos over raw float bit patterns [lo_bits, hi_bits) with the given stride (and
// the negated arguments): out[0] = results that differ from the CORRECTLY ROUNDED value (double libm rounded to
// float), out[1] = results that differ from libm sinf/cosf, out[2] = the largest such difference in ulps.
void hko_trig_check(uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, int64_t* out) {
  out[0] = out[1] = out[2] = 0;
  for (uint64_t u = lo_bits; u < hi_bits; u += stride) {
    for (int sgn = 0; sgn < 2; ++sgn) {
      float x = u2f((uint32_t)u);
      if (sgn) x = -x;
      double sd, cd;
      sincos_poly((double)x, &sd, &cd);
      float got[2] = {(float)sd, (float)cd};
      float cr[2] = {(float)std::sin((double)x), (float)std::cos((double)x)};
      float lm[2] = {sinf(x), cosf(x)};
      for (int k = 0; k < 2; ++k) {
        if (got[k] != cr[k]) out[0]++;
        if (got[k] != lm[k]) {
          out[1]++;
          int64_t d = (int64_t)f2u(got[k]) - (int64_t)f2u(lm[k]);
          if (d < 0) d = -d;
          if (d > out[2]) out[2] = d;
        }
      }
    }
  }
}

}  // extern "C"
