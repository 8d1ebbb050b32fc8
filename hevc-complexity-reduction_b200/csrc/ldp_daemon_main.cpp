// resi_to_cu_depth_LDP -- drop-in for `python resi_to_cu_depth_LDP.py` of the inter-mode (LDP) encoder
// (HM-16.5_Test_LDP/bin/resi_to_cu_depth_LDP.py:146-187, protocol in README.md:64-84): started by hand next to the
// encoder, it serves the file-signal handshake from the current directory until killed.
//
// Environment: ETHCNN_MODEL_DIR (default "."), ETHCNN_DAEMON_MAX_FRAMES and ETHCNN_DAEMON_IDLE_MS (default 0 = forever).
#include <cstdio>
#include <cstdlib>

#include "../../include/ethcnn.h"

int main(int argc, char** argv) {
  const char* dir = argc > 1 ? argv[1] : ".";
  const char* model_dir = getenv("ETHCNN_MODEL_DIR");
  if (!model_dir || !*model_dir) model_dir = dir;
  const int max_frames = getenv("ETHCNN_DAEMON_MAX_FRAMES") ? atoi(getenv("ETHCNN_DAEMON_MAX_FRAMES")) : 0;
  const int idle_ms = getenv("ETHCNN_DAEMON_IDLE_MS") ? atoi(getenv("ETHCNN_DAEMON_IDLE_MS")) : 0;
  ethcnn_handle* h = nullptr;
  if (ethcnn_create(model_dir, nullptr, ETHCNN_MODE_LDP, 1, &h) != ETHCNN_OK) {
    fprintf(stderr, "resi_to_cu_depth_LDP: %s\n", ethcnn_last_error());
    return 1;
  }
  printf("ethcnn: predictor initialized.\n");  // the reference prints 'Python: Tensorflow initialized.' here
  fflush(stdout);
  const int rc = ethcnn_ldp_serve(h, dir, max_frames, idle_ms);
  if (rc < 0) fprintf(stderr, "resi_to_cu_depth_LDP: %s\n", ethcnn_last_error());
  else printf("%d frames predicted.\n", rc);
  ethcnn_destroy(h);
  return rc < 0 ? 1 : 0;
}
