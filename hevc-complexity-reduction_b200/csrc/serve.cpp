// Resident server for the All-Intra drop-in (ethcnn_serve / ethcnn_request, include/ethcnn.h).
//
// Why: the patched HM encoder starts `python video_to_cu_depth.py <yuv> <W> <H> <QP>` afresh for every encode
// (TAppEncCfg.cpp:2317-2321) and a fresh process pays for CUDA initialisation before it can do anything: 2 - 4 s on a B200
// box (driver init + primary context; tools/cli_trace.sh) against ~0.15 s of actual work for a 50-frame 4928x3264 file.
// The reference has the same problem with TensorFlow's start-up and solves it for the inter-mode path with a resident
// daemon (README.md:64-84); this is the same idea for the intra path.  A server process keeps the context, the packed
// weights and the scratch buffers alive and serves requests over a Unix-domain stream socket; the drop-in (CLI or Python
// shim) becomes a client when ETHCNN_SERVER names the socket, and falls back to working in-process when nobody listens.
//
// Protocol (one request per connection, text, '\n' terminated, fields separated by '\t'):
//   request   "PREDICT\t<cwd>\t<yuv>\t<width>\t<height>\t<qp>\t<out>\n"     relative paths are resolved against <cwd>
//   reply     "0\n" on success, "<negative ETHCNN_E_* code>\t<message>\n" otherwise
//   request   "QUIT\n" makes the server return (reply "0\n")
// Every PREDICT request is answered through ethcnn_predict_yuv_file_from: Thr_info.txt is re-read from the CLIENT's directory
// and the client's checkpoint must be the resident one (the reference re-reads both on every run, net_CNN.py:47,
// video_to_cu_depth.py:126-133) -- a stale threshold or a different model is an error, never a silent answer.
// Trust: the socket is created 0600, peers are checked with SO_PEERCRED (same uid or root), a connection that sends nothing
// is dropped after 5 s, and a second server refuses to start on a socket that a live server still answers.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE   // struct ucred
#endif
#include <poll.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ethcnn.h"

namespace {

thread_local std::string g_serve_error;

bool fill_addr(const char* path, sockaddr_un* addr) {
  memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  if (!path || !*path || strlen(path) >= sizeof(addr->sun_path)) return false;
  strcpy(addr->sun_path, path);
  return true;
}

bool write_all(int fd, const std::string& s) {
  size_t done = 0;
  while (done < s.size()) {
    const ssize_t n = send(fd, s.data() + done, s.size() - done, MSG_NOSIGNAL);   // a vanished peer must not kill the server
    if (n <= 0) {
      if (n < 0 && errno == EINTR) continue;
      return false;
    }
    done += size_t(n);
  }
  return true;
}

// one '\n'-terminated line (without the terminator); false on EOF / error / oversize
bool read_line(int fd, std::string* line) {
  line->clear();
  char c;
  while (line->size() < 16384) {
    const ssize_t n = read(fd, &c, 1);
    if (n == 0) return false;
    if (n < 0) {
      if (errno == EINTR) continue;
      return false;   // includes EAGAIN: the receive timeout of an accepted socket expired
    }
    if (c == '\n') return true;
    line->push_back(c);
  }
  return false;
}

std::vector<std::string> split_tabs(const std::string& s) {
  std::vector<std::string> out;
  size_t a = 0;
  for (;;) {
    const size_t b = s.find('\t', a);
    out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  return out;
}

std::string resolve(const std::string& cwd, const std::string& p) { return (!p.empty() && p[0] == '/') ? p : cwd + "/" + p; }

bool to_int(const std::string& s, int* v) {
  char* end = nullptr;
  const long x = strtol(s.c_str(), &end, 10);
  if (s.empty() || *end) return false;
  *v = int(x);
  return true;
}

}  // namespace

extern "C" {

int ethcnn_serve(ethcnn_handle* h, const char* socket_path, int max_requests, int idle_timeout_ms) {
  if (!h) return ETHCNN_E_ARG;
  sockaddr_un addr;
  if (!fill_addr(socket_path, &addr)) return ETHCNN_E_ARG;
  {   // is a live server already answering on this path?  Then do not orphan it.
    const int probe = socket(AF_UNIX, SOCK_STREAM, 0);
    if (probe >= 0) {
      const bool alive = connect(probe, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) == 0;
      close(probe);
      if (alive) {
        g_serve_error = std::string("a live server already listens on ") + socket_path;
        return ETHCNN_E_IO;
      }
    }
  }
  const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
  if (ls < 0) return ETHCNN_E_IO;
  unlink(socket_path);   // a stale socket of a dead server (nobody answered the probe above)
  const mode_t old_mask = umask(0077);   // the socket file is created rw for the owner only
  const bool bound = bind(ls, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) == 0;
  umask(old_mask);
  if (!bound || chmod(socket_path, 0600) != 0 || listen(ls, 16) != 0) {
    close(ls);
    return ETHCNN_E_IO;
  }
  int served = 0;
  bool quit = false;
  while (!quit && (max_requests <= 0 || served < max_requests)) {
    pollfd pfd{ls, POLLIN, 0};
    const int pr = poll(&pfd, 1, idle_timeout_ms > 0 ? idle_timeout_ms : -1);
    if (pr == 0) break;   // idle
    if (pr < 0) {
      if (errno == EINTR) continue;
      break;
    }
    const int cs = accept(ls, nullptr, nullptr);
    if (cs < 0) continue;
    {   // a client that connects and stays silent must not block everybody else (QUIT included)
      timeval tv{5, 0};
      setsockopt(cs, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
      setsockopt(cs, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    }
    ucred cred{};
    socklen_t cl = sizeof(cred);
    if (getsockopt(cs, SOL_SOCKET, SO_PEERCRED, &cred, &cl) != 0 || (cred.uid != geteuid() && cred.uid != 0)) {
      write_all(cs, std::to_string(ETHCNN_E_ARG) + "\tpeer uid is not the server's\n");
      close(cs);
      continue;
    }
    std::string line, reply;
    if (read_line(cs, &line)) {
      const std::vector<std::string> f = split_tabs(line);
      int w = 0, hgt = 0, qp = 0;
      if (f.size() == 1 && f[0] == "QUIT") {
        quit = true;
        reply = "0\n";
      } else if (f.size() == 7 && f[0] == "PREDICT" && to_int(f[3], &w) && to_int(f[4], &hgt) && to_int(f[5], &qp)) {
        const int rc = ethcnn_predict_yuv_file_from(h, f[1].c_str(), resolve(f[1], f[2]).c_str(), w, hgt, qp, resolve(f[1], f[6]).c_str());
        reply = rc == ETHCNN_OK ? std::string("0\n") : std::to_string(rc) + "\t" + ethcnn_last_error() + "\n";
        ++served;
      } else {
        reply = std::to_string(ETHCNN_E_ARG) + "\tmalformed request\n";
      }
      write_all(cs, reply);
    }
    close(cs);
  }
  close(ls);
  unlink(socket_path);
  return served;
}

int ethcnn_request(const char* socket_path, const char* yuv_path, int width, int height, int qp, const char* out_path) {
  sockaddr_un addr;
  if (!fill_addr(socket_path, &addr) || !yuv_path || !out_path) return ETHCNN_E_ARG;
  const int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) return ETHCNN_E_IO;
  if (connect(s, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
    close(s);
    g_serve_error = std::string("no server listens on ") + socket_path;
    return ETHCNN_E_IO;
  }
  char cwd[4096];
  if (!getcwd(cwd, sizeof(cwd))) strcpy(cwd, ".");
  const std::string req = std::string("PREDICT\t") + cwd + "\t" + yuv_path + "\t" + std::to_string(width) + "\t" + std::to_string(height) +
                          "\t" + std::to_string(qp) + "\t" + out_path + "\n";
  std::string line;
  int rc = ETHCNN_E_IO;
  if (write_all(s, req) && read_line(s, &line)) {
    const std::vector<std::string> f = split_tabs(line);
    if (!to_int(f[0], &rc)) rc = ETHCNN_E_IO;
    g_serve_error = f.size() > 1 ? f[1] : std::string();
  } else {
    g_serve_error = "the server closed the connection";
  }
  close(s);
  return rc;
}

int ethcnn_request_quit(const char* socket_path) {
  sockaddr_un addr;
  if (!fill_addr(socket_path, &addr)) return ETHCNN_E_ARG;
  const int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) return ETHCNN_E_IO;
  if (connect(s, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
    close(s);
    return ETHCNN_E_IO;
  }
  std::string line;
  const bool ok = write_all(s, "QUIT\n") && read_line(s, &line);
  close(s);
  return ok ? ETHCNN_OK : ETHCNN_E_IO;
}

const char* ethcnn_request_error(void) { return g_serve_error.c_str(); }

}  // extern "C"
