// wire.cpp -- the reference's ASCII wire format and an asynchronous state feed, OFF the hot path.
//
// The reference blocks its control loop on the serial port (src/solver.cpp:43-74: readPort -> controllerStep -> writePort).
// For a batch of controllers one slow device must not stall the others: a reader thread per device parses frames as
// SerialPort::getDataFromSerial does (src/SerialPort.cpp:106-138) and keeps the most recent state; the batch driver polls
// it without blocking and copies the state into its lane of X before smpc_mpc_set_state.
// Deliberate fixes ("belong here, not on the hot path"): dt is returned (readPort takes it by value,
// SerialPort.cpp:142, so the reference always sees 0); writePort's length sizeof(char*) (cpp:165) is a parameter.
#include <poll.h>
#include <pthread.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/solvempc_b200.h"
#include "handles.hpp"

namespace {
constexpr int kFrameBuf = 42;      // char read_buf[42], include/SerialPort.h:82
constexpr int kMinFrame = 30;      // readPort accepts num_bytes > 30, src/SerialPort.cpp:146
}  // namespace

struct smpc_feed {
  int fd = -1;
  pthread_t thread{};
  bool started = false;
  std::atomic<bool> stop{false};
  std::mutex mu;
  double dt = 0.0, X[4] = {0, 0, 0, 0};
  long long seq = 0, accepted = 0, rejected = 0;
};

// length_checked != 0: the caller (the reader thread) has already applied readPort's byte-count test to the raw read
static int parse_frame_impl(const char *buf, int nbytes, double *dt, double *X, int length_checked) {
  if (!buf || nbytes <= 0 || (!length_checked && nbytes <= kMinFrame)) return 0;   // cpp:146
  // getDataFromSerial: strtok on " ", atof, stored through float ref[5]; a field that parses to 0 leaves the 0 default
  char tmp[256];
  const int len = nbytes < (int)sizeof(tmp) - 1 ? nbytes : (int)sizeof(tmp) - 1;
  std::memcpy(tmp, buf, len);
  tmp[len] = '\0';
  float ref[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  char *save = nullptr;
  char *ptr = strtok_r(tmp, " ", &save);
  for (int index = 0; index < 5; ++index) {
    const double temp = ptr ? std::atof(ptr) : 0.0;               // (the reference would dereference NULL here)
    if (temp != 0) ref[index] = (float)temp;
    ptr = ptr ? strtok_r(nullptr, " ", &save) : nullptr;
  }
  if (dt) *dt = ref[0];
  if (X) { X[0] = ref[1]; X[1] = ref[2]; X[2] = ref[3]; X[3] = ref[4]; }
  return 1;
}

extern "C" {

int smpc_wire_parse_frame(const char *buf, int nbytes, double *dt, double *X) { return parse_frame_impl(buf, nbytes, dt, X, 0); }

int smpc_wire_format_control(double U, char *out, int capacity, int max_chars) {
  if (!out || capacity <= 0) return 0;
  const std::string s = std::to_string(U);                         // cpp:165
  int n = (int)s.size();
  if (max_chars > 0 && n > max_chars) n = max_chars;               // the reference sends sizeof(char*) = 8 characters
  if (n > capacity - 1) n = capacity - 1;
  std::memcpy(out, s.data(), n);
  out[n] = '\0';
  return n;
}

static void *feed_main(void *arg) {
  smpc_feed *f = static_cast<smpc_feed *>(arg);
  char line[kFrameBuf + 1];
  int fill = 0;
  auto flush = [&](int terminator) {
    double dt, X[4];
    // readPort tests the raw byte count of read(), the terminator included (src/SerialPort.cpp:146: num_bytes > 30)
    const int ok = (fill + terminator > kMinFrame) ? parse_frame_impl(line, fill, &dt, X, 1) : 0;
    std::lock_guard<std::mutex> g(f->mu);
    if (ok) { f->dt = dt; std::memcpy(f->X, X, sizeof(X)); ++f->seq; ++f->accepted; }
    else if (fill > 0) ++f->rejected;
    fill = 0;
  };
  while (!f->stop.load()) {
    struct pollfd p = {f->fd, POLLIN, 0};
    const int r = poll(&p, 1, 50);
    if (r <= 0) continue;
    if (p.revents & (POLLERR | POLLNVAL)) break;
    char chunk[64];
    const ssize_t got = read(f->fd, chunk, sizeof(chunk));
    if (got <= 0) { if (p.revents & POLLHUP) break; continue; }
    for (ssize_t k = 0; k < got; ++k) {
      // a frame ends at a newline or when the reference's 42-byte buffer is full
      if (chunk[k] == '\n') { flush(1); continue; }
      line[fill++] = chunk[k];
      if (fill == kFrameBuf) flush(0);
    }
  }
  return nullptr;
}

int smpc_feed_open(smpc_feed **out, int fd) {
  if (!out) return smpc::fail(SMPC_ERR_ARG, "out is null");
  *out = nullptr;
  if (fd < 0) return smpc::fail(SMPC_ERR_ARG, "bad file descriptor");
  smpc_feed *f = new smpc_feed;
  f->fd = fd;
  if (pthread_create(&f->thread, nullptr, feed_main, f) != 0) { delete f; return smpc::fail(SMPC_ERR_IO, "cannot start the reader thread"); }
  f->started = true;
  *out = f;
  return SMPC_OK;
}

int smpc_feed_latest(smpc_feed *f, long long *seq, double *dt, double *X) {
  if (!f || !seq) return 0;
  std::lock_guard<std::mutex> g(f->mu);
  if (f->seq == *seq) return 0;
  *seq = f->seq;
  if (dt) *dt = f->dt;
  if (X) std::memcpy(X, f->X, sizeof(f->X));
  return 1;
}

int smpc_feed_stats(smpc_feed *f, long long *accepted, long long *rejected) {
  if (!f) return smpc::fail(SMPC_ERR_ARG, "null handle");
  std::lock_guard<std::mutex> g(f->mu);
  if (accepted) *accepted = f->accepted;
  if (rejected) *rejected = f->rejected;
  return SMPC_OK;
}

int smpc_feed_close(smpc_feed *f) {
  if (!f) return SMPC_OK;
  f->stop.store(true);
  if (f->started) pthread_join(f->thread, nullptr);
  delete f;
  return SMPC_OK;
}

}  // extern "C"
