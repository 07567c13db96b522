// orb_io.cu — the fork's binary dump formats for keypoints and descriptors, host-side file I/O only:
//   descriptors  include/SaveLoadWorld.h:1448-1459   per keyframe { 0xEB 0x90, int32 n, n x 32 bytes }
//   keypoints    include/SaveLoadWorld.h:1408-1424   per keyframe { 0xEB 0x90, size_t n, n x (pt.x pt.y size angle response : f32, octave class_id : i32) }
// A dump of the reference's map is therefore directly usable as the descriptor database of orb_hamming_knn2 (config 5): the rows
// of all records are concatenated, rec_start[] keeps the keyframe boundaries.
#include "orb_internal.h"
#include <cstdio>
#include <cstring>
#include <climits>
#include <sys/types.h>

namespace {

template <typename CountT, size_t ROW>
int read_records(const char* path, uint8_t* rows, int64_t cap_rows, int32_t* rec_start, int cap_records, int64_t* nrows, int32_t* nrecords)
{
    if (!path || !nrows || !nrecords) return ORB_ERR_INVALID;
    *nrows = 0; *nrecords = 0;
    FILE* f = fopen(path, "rb");
    if (!f) return ORB_ERR_INVALID;
    // a record count is trusted only as far as the file can hold it
    if (fseeko(f, 0, SEEK_END) != 0) { fclose(f); return ORB_ERR_INVALID; }
    const int64_t fsize = (int64_t)ftello(f);
    rewind(f);
    int64_t total = 0;
    int32_t recs = 0;
    int rc = ORB_OK;
    for (;;) {
        unsigned char hdr[2];
        if (fread(hdr, 1, 2, f) != 2) break;                           // clean end of file
        CountT n = 0;
        if (hdr[0] != 0xEB || hdr[1] != 0x90 || fread(&n, sizeof n, 1, f) != 1 || (int64_t)n < 0) { rc = ORB_ERR_INVALID; break; }
        const int64_t here = (int64_t)ftello(f);
        if (here < 0 || (uint64_t)n > (uint64_t)(fsize - here) / ROW) { rc = ORB_ERR_INVALID; break; }      // truncated / corrupt record
        if (total + (int64_t)n > INT32_MAX || recs == INT32_MAX) { rc = ORB_ERR_CAPACITY; break; }            // row offsets are int32 (rec_start, match indices)
        if (rec_start && recs < cap_records) rec_start[recs] = (int32_t)total;
        if (rows && total + (int64_t)n <= cap_rows) {
            if (n && fread(rows + (size_t)total * ROW, ROW, (size_t)n, f) != (size_t)n) { rc = ORB_ERR_INVALID; break; }
        } else if (fseeko(f, (off_t)((uint64_t)n * ROW), SEEK_CUR) != 0) { rc = ORB_ERR_INVALID; break; }
        total += (int64_t)n;
        recs++;
    }
    fclose(f);
    if (rc != ORB_OK) return rc;
    if (rec_start && recs < cap_records) rec_start[recs] = (int32_t)total;      // closing offset when there is room
    *nrows = total; *nrecords = recs;
    if ((rows && total > cap_rows) || (rec_start && recs + 1 > cap_records)) return ORB_ERR_CAPACITY;
    return ORB_OK;
}

template <typename CountT, size_t ROW>
int write_records(const char* path, const uint8_t* rows, const int32_t* rec_start, int nrecords)
{
    if (!path || nrecords < 0 || (nrecords > 0 && !rec_start)) return ORB_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return ORB_ERR_INVALID;
    const unsigned char hdr[2] = { 0xEB, 0x90 };
    int rc = ORB_OK;
    for (int r = 0; r < nrecords && rc == ORB_OK; r++) {
        const int64_t n64 = (int64_t)rec_start[r + 1] - rec_start[r];
        if (n64 < 0 || (n64 > 0 && !rows)) { rc = ORB_ERR_INVALID; break; }
        const CountT n = (CountT)n64;
        if (fwrite(hdr, 1, 2, f) != 2 || fwrite(&n, sizeof n, 1, f) != 1 ||
            (n64 && fwrite(rows + (size_t)rec_start[r] * ROW, ROW, (size_t)n64, f) != (size_t)n64)) rc = ORB_ERR_INVALID;
    }
    if (fclose(f) != 0) rc = ORB_ERR_INVALID;
    return rc;
}

} // namespace

extern "C" {

int orb_db_read_descriptors(const char* path, uint8_t* desc, int64_t cap_rows, int32_t* rec_start, int cap_records, int64_t* nrows,
                            int32_t* nrecords)
{
    return read_records<int32_t, 32>(path, desc, cap_rows, rec_start, cap_records, nrows, nrecords);
}
int orb_db_write_descriptors(const char* path, const uint8_t* desc, const int32_t* rec_start, int nrecords)
{
    return write_records<int32_t, 32>(path, desc, rec_start, nrecords);
}
int orb_db_read_keypoints(const char* path, orb_keypoint* kps, int64_t cap_rows, int32_t* rec_start, int cap_records, int64_t* nrows,
                          int32_t* nrecords)
{
    static_assert(sizeof(orb_keypoint) == 28, "keypoint record is 28 bytes");
    return read_records<uint64_t, 28>(path, reinterpret_cast<uint8_t*>(kps), cap_rows, rec_start, cap_records, nrows, nrecords);
}
int orb_db_write_keypoints(const char* path, const orb_keypoint* kps, const int32_t* rec_start, int nrecords)
{
    return write_records<uint64_t, 28>(path, reinterpret_cast<const uint8_t*>(kps), rec_start, nrecords);
}

} // extern "C"
