// video_to_cu_depth -- drop-in for `python video_to_cu_depth.py <yuv> <W> <H> <QP>`
// (HM-16.5_Test_AI/bin/video_to_cu_depth.py:120-145; invoked by the patched HM encoder through
// system(), TAppEncCfg.cpp:2317-2321).  Same argv, same files read from the cwd (Thr_info.txt,
// model_2000000_qp*.dat.*), same cu_depth.dat written to the cwd, exit status 0 on success.
//
// Environment: ETHCNN_GPUS = number of GPUs to shard frames across (default 1);
//              ETHCNN_MODEL_DIR = directory of the checkpoints / Thr_info.txt (default ".");
//              ETHCNN_SERVER = Unix socket of a resident server (`video_to_cu_depth --serve <socket>`): the request is
//              handed to it (no CUDA start-up in this process); if nobody listens the work is done in-process.
//
// `video_to_cu_depth --serve <socket> [idle_timeout_ms]` runs that server in the foreground (checkpoints and Thr_info.txt
// from ETHCNN_MODEL_DIR or the cwd) until `video_to_cu_depth --quit <socket>`, the idle timeout or a signal.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/ethcnn.h"

static int serve_main(int argc, char** argv) {
  const char* model_dir = getenv("ETHCNN_MODEL_DIR");
  if (!model_dir || !*model_dir) model_dir = ".";
  int n_gpus = 1;
  if (const char* g = getenv("ETHCNN_GPUS")) n_gpus = atoi(g) > 0 ? atoi(g) : 1;
  ethcnn_handle* h = nullptr;
  if (ethcnn_create(model_dir, nullptr, ETHCNN_MODE_AI, n_gpus, &h) != ETHCNN_OK) {
    fprintf(stderr, "video_to_cu_depth: %s\n", ethcnn_last_error());
    return 1;
  }
  fprintf(stderr, "video_to_cu_depth: serving on %s\n", argv[2]);
  const int served = ethcnn_serve(h, argv[2], 0, argc > 3 ? atoi(argv[3]) : 0);
  ethcnn_destroy(h);
  if (served < 0) return fprintf(stderr, "video_to_cu_depth: cannot serve on %s\n", argv[2]), 1;
  fprintf(stderr, "video_to_cu_depth: served %d requests\n", served);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && strcmp(argv[1], "--serve") == 0) return serve_main(argc, argv);
  if (argc == 3 && strcmp(argv[1], "--quit") == 0) return ethcnn_request_quit(argv[2]) == ETHCNN_OK ? 0 : 1;
  if (argc != 5) {  // python: assert len(sys.argv) == 5
    fprintf(stderr, "usage: %s <yuv_file> <width> <height> <qp>\n", argv[0]);
    return 1;
  }
  const char* yuv = argv[1];
  char* end = nullptr;
  const long width = strtol(argv[2], &end, 10);
  if (*end || end == argv[2]) return fprintf(stderr, "invalid width '%s'\n", argv[2]), 1;
  const long height = strtol(argv[3], &end, 10);
  if (*end || end == argv[3]) return fprintf(stderr, "invalid height '%s'\n", argv[3]), 1;
  const long qp = strtol(argv[4], &end, 10);
  if (*end || end == argv[4]) return fprintf(stderr, "invalid qp '%s'\n", argv[4]), 1;

  const char* model_dir = getenv("ETHCNN_MODEL_DIR");
  if (!model_dir || !*model_dir) model_dir = ".";
  int n_gpus = 1;
  if (const char* g = getenv("ETHCNN_GPUS")) n_gpus = atoi(g) > 0 ? atoi(g) : 1;

  if (const char* srv = getenv("ETHCNN_SERVER")) {
    if (*srv) {
      const auto t1 = std::chrono::steady_clock::now();
      const int rc = ethcnn_request(srv, yuv, int(width), int(height), int(qp), "cu_depth.dat");
      if (rc == ETHCNN_OK) {
        printf("--------\n\nPredicting Time: %.3f sec.\n\n--------\n",
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
        return 0;
      }
      // two cases are not failures of this run: nobody listens, or the server holds OTHER weights than this directory's
      // checkpoint (it refuses rather than answer with them) -- both are served in-process, like a run without a server
      const bool nobody = rc == ETHCNN_E_IO && strstr(ethcnn_request_error(), "no server") != nullptr;
      const bool other_weights = rc == ETHCNN_E_FORMAT && strstr(ethcnn_request_error(), "resident handle") != nullptr;
      if (!nobody && !other_weights) {
        fprintf(stderr, "video_to_cu_depth: server: %s\n", ethcnn_request_error());
        return 1;
      }
      fprintf(stderr, "video_to_cu_depth: %s, working in-process\n", ethcnn_request_error());
    }
  }
  ethcnn_handle* h = nullptr;
  int rc = ethcnn_create(model_dir, nullptr, ETHCNN_MODE_AI, n_gpus, &h);
  if (rc != ETHCNN_OK) {
    fprintf(stderr, "video_to_cu_depth: %s\n", ethcnn_last_error());
    return 1;
  }
  const auto t1 = std::chrono::steady_clock::now();
  rc = ethcnn_predict_yuv_file(h, yuv, int(width), int(height), int(qp), "cu_depth.dat");
  const auto t2 = std::chrono::steady_clock::now();
  if (rc != ETHCNN_OK) {
    fprintf(stderr, "video_to_cu_depth: %s\n", ethcnn_last_error());
    ethcnn_destroy(h);
    return 1;
  }
  ethcnn_destroy(h);
  // same closing banner as video_to_cu_depth.py:145 (HM ignores stdout)
  printf("--------\n\nPredicting Time: %.3f sec.\n\n--------\n", std::chrono::duration<double>(t2 - t1).count());
  return 0;
}
