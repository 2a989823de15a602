/* Plain C99 client of the C-ABI (no Python, no torch, no CUDA headers): links libcp360.so, calls the host-only entry
 * points and checks the worked example of SURVEY.md section 8c (CubePad(1) of a 6x1x4x4 arange, face 0 = Back and
 * face 5 = Top) plus the error conventions. Built and run by tests/test_c_abi_client.py. */
#include <stdio.h>
#include <string.h>

#include "cp360.h"

static const int kBack[36] = {83, 83, 82, 81, 80, 80, 67, 0, 1, 2, 3, 48, 71, 4, 5, 6, 7, 52,
                              75, 8, 9, 10, 11, 56, 79, 12, 13, 14, 15, 60, 31, 31, 30, 29, 28, 28};
static const int kTop[36] = {3, 3, 2, 1, 0, 0, 48, 80, 81, 82, 83, 67, 49, 84, 85, 86, 87, 66,
                             50, 88, 89, 90, 91, 65, 51, 92, 93, 94, 95, 64, 32, 32, 33, 34, 35, 35};

int main(void) {
  int32_t map[6 * 36];
  int ho = 0, wo = 0, i;
  if (cp360_version() != CP360_VERSION) return 1;
  if (cp360_cubepad_out_shape(4, 4, 1, 1, 1, 1, &ho, &wo) != CP360_OK || ho != 6 || wo != 6) return 2;
  if (cp360_cubepad_build_map(4, 4, 1, 1, 1, 1, map) != CP360_OK) return 3;
  for (i = 0; i < 36; ++i) {
    if (map[i] != kBack[i]) return 4;
    if (map[5 * 36 + i] != kTop[i]) return 5;
  }
  /* batch not a multiple of 6: the reference prints 'CubePad size mismatch!' and exits (cube_pad.py:33-35) */
  if (cp360_cubepad_fwd(NULL, NULL, 5, 1, 4, 4, 1, 1, 1, 1, 4, NULL) != CP360_ERR_GROUP) return 6;
  if (strstr(cp360_last_error(), "size mismatch") == NULL) return 7;
  if (cp360_cubepad_out_shape(4, 5, 1, 1, 1, 1, &ho, &wo) != CP360_ERR_SHAPE) return 8;   /* H != W */
  if (cp360_e2c_build_map(8, 32, 60, 90.0, NULL, NULL, NULL, NULL, NULL) != CP360_ERR_SHAPE) return 9; /* W != 2H */
  printf("c-abi host client ok: version %d, %s\n", cp360_version(), cp360_status_string(CP360_OK));
  return 0;
}
