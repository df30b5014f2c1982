/* TEST INFRASTRUCTURE ONLY -- never shipped, never loaded by the product.
 *
 * A stand-in for the eight NCCL entry points that csrc/tu_shard.cu binds at run time
 * (ncclGetUniqueId, ncclCommInitRank, ncclCommDestroy, ncclGroupStart / End, ncclSend, ncclRecv,
 * ncclGetErrorString), so that the element-sharded flow of the library (sse_shard_create /
 * sse_shard_residual / sse_shard_rk_step_ck54: partition, pack, exchange, interior and boundary
 * ranges on their streams, unpack) can run on the CPU emulator (tests/emu) with one PROCESS per
 * rank -- tests/test_sharded_library_cpu.py.  Built as libnccl.so.2 (soname) and loaded into the
 * worker process before the emulation library dlopen()s "libnccl.so.2".
 *
 * Transport: one file per message in the directory named by the 128-byte "unique id",
 * msg_<src>_<dst>_<seq>.bin, written under a temporary name and renamed (atomic).  "Device"
 * pointers of the emulator are host pointers.  Sends complete at once; receives posted inside a
 * group are deferred to ncclGroupEnd, so the recv-then-send order of tu_shard.cu cannot deadlock;
 * a receive polls for its file for at most FAKE_NCCL_TIMEOUT_S (default 120) seconds.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <sys/stat.h>

typedef struct { char internal[128]; } ncclUniqueId;

typedef struct {
  char dir[128];
  int rank, world;
  long send_seq[64], recv_seq[64];
} fake_comm;

typedef struct { void* buf; size_t bytes; int peer; fake_comm* c; } pending_recv;

static int g_group = 0;
static pending_recv g_pending[256];
static int g_npending = 0;
static char g_err[256] = "ok";

static size_t type_size(int t) { return (t == 8 || t == 4 || t == 5) ? 8 : ((t == 7 || t == 2 || t == 3) ? 4 : 1); }

static int do_recv(pending_recv* p) {
  char path[320];
  fake_comm* c = p->c;
  snprintf(path, sizeof path, "%s/msg_%d_%d_%ld.bin", c->dir, p->peer, c->rank, c->recv_seq[p->peer]);
  const char* te = getenv("FAKE_NCCL_TIMEOUT_S");
  const double limit = te ? atof(te) : 120.0;
  struct timespec t0, t1, nap = {0, 1000000};
  clock_gettime(CLOCK_MONOTONIC, &t0);
  FILE* f = NULL;
  for (;;) {
    f = fopen(path, "rb");
    if (f) break;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > limit) {
      snprintf(g_err, sizeof g_err, "fake nccl: rank %d timed out waiting for %s", c->rank, path);
      return 1;
    }
    nanosleep(&nap, NULL);
  }
  size_t got = fread(p->buf, 1, p->bytes, f);
  fseek(f, 0, SEEK_END);
  long total = ftell(f);
  fclose(f);
  if (got != p->bytes || (size_t)total != p->bytes) {
    snprintf(g_err, sizeof g_err, "fake nccl: %s holds %ld bytes, receive wants %zu", path, total, p->bytes);
    return 2;
  }
  unlink(path);
  c->recv_seq[p->peer]++;
  return 0;
}

int ncclGetUniqueId(ncclUniqueId* id) {
  memset(id, 0, sizeof *id);
  char tmpl[] = "/tmp/fake_nccl_XXXXXX";
  if (!mkdtemp(tmpl)) { snprintf(g_err, sizeof g_err, "fake nccl: mkdtemp failed"); return 1; }
  strncpy(id->internal, tmpl, sizeof id->internal - 1);
  return 0;
}

int ncclCommInitRank(void** comm, int world, ncclUniqueId id, int rank) {
  if (world > 64 || rank < 0 || rank >= world) { snprintf(g_err, sizeof g_err, "fake nccl: bad rank / world"); return 1; }
  fake_comm* c = (fake_comm*)calloc(1, sizeof *c);
  memcpy(c->dir, id.internal, sizeof c->dir);
  c->dir[sizeof c->dir - 1] = 0;
  struct stat st;
  if (stat(c->dir, &st) != 0 || !S_ISDIR(st.st_mode)) {
    snprintf(g_err, sizeof g_err, "fake nccl: id does not name a directory: %s", c->dir);
    free(c);
    return 1;
  }
  c->rank = rank;
  c->world = world;
  *comm = c;
  return 0;
}

int ncclCommDestroy(void* comm) { free(comm); return 0; }

int ncclGroupStart(void) { g_group++; return 0; }

int ncclGroupEnd(void) {
  if (--g_group > 0) return 0;
  int rc = 0;
  for (int i = 0; i < g_npending && !rc; ++i) rc = do_recv(&g_pending[i]);
  g_npending = 0;
  return rc;
}

int ncclSend(const void* buf, size_t count, int type, int peer, void* comm, void* stream) {
  (void)stream;
  fake_comm* c = (fake_comm*)comm;
  char tmp[320], path[320];
  snprintf(path, sizeof path, "%s/msg_%d_%d_%ld.bin", c->dir, c->rank, peer, c->send_seq[peer]);
  snprintf(tmp, sizeof tmp, "%s/tmp_%d_%d_%ld", c->dir, c->rank, peer, c->send_seq[peer]);
  FILE* f = fopen(tmp, "wb");
  if (!f) { snprintf(g_err, sizeof g_err, "fake nccl: cannot write %s", tmp); return 1; }
  const size_t bytes = count * type_size(type);
  if (fwrite(buf, 1, bytes, f) != bytes) { fclose(f); snprintf(g_err, sizeof g_err, "fake nccl: short write"); return 1; }
  fclose(f);
  if (rename(tmp, path) != 0) { snprintf(g_err, sizeof g_err, "fake nccl: rename failed"); return 1; }
  c->send_seq[peer]++;
  return 0;
}

int ncclRecv(void* buf, size_t count, int type, int peer, void* comm, void* stream) {
  (void)stream;
  pending_recv p = {buf, count * type_size(type), peer, (fake_comm*)comm};
  if (g_group > 0) {
    if (g_npending >= 256) { snprintf(g_err, sizeof g_err, "fake nccl: too many receives in a group"); return 1; }
    g_pending[g_npending++] = p;
    return 0;
  }
  return do_recv(&p);
}

const char* ncclGetErrorString(int rc) { (void)rc; return g_err; }
