/* fm_blob.h — layout of the "model blob" passed to fm_create() (and to the test oracle).
 *
 * The blob is produced by factory_marl_b200/model.py:pack_blob() from the scene constants of the
 * reference (challenge_env/challenge_env/scene.py:109-169 + assets XML files).  It is a tagged list of
 * little-endian arrays:
 *
 *   uint32 magic ('FMB1' = 0x464D4231), uint32 n_entries,
 *   n_entries x fm_blob_entry, padding to 16 bytes, payload (each array 16-byte aligned).
 *
 * This header is data-format only (no algorithm code); it is shared by the CUDA library and by the
 * CPU oracle so that both read the same constants.
 */
#ifndef FM_BLOB_H
#define FM_BLOB_H
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define FM_BLOB_MAGIC 0x464D4231u

typedef struct fm_blob_entry {
  char name[32];
  uint32_t dtype;    /* 0 = int32, 1 = float64 */
  uint32_t ndim;
  uint32_t shape[4];
  uint32_t offset;   /* bytes from blob start */
  uint32_t pad;
} fm_blob_entry;

/* returns pointer to array data or NULL; count (number of elements) optional */
static inline const void* fm_blob_find(const void* blob, size_t nbytes, const char* name, uint32_t dtype,
                                       size_t* count) {
  const uint32_t* h = (const uint32_t*)blob;
  if (nbytes < 8 || h[0] != FM_BLOB_MAGIC) return NULL;
  uint32_t n = h[1];
  const fm_blob_entry* e = (const fm_blob_entry*)((const char*)blob + 8);
  for (uint32_t i = 0; i < n; ++i) {
    if (strncmp(e[i].name, name, 32) == 0) {
      if (e[i].dtype != dtype) return NULL;
      size_t c = 1;
      for (uint32_t d = 0; d < 4; ++d) c *= e[i].shape[d];
      if (count) *count = c;
      if (e[i].offset + c * (dtype ? 8 : 4) > nbytes) return NULL;
      return (const char*)blob + e[i].offset;
    }
  }
  return NULL;
}
#endif
