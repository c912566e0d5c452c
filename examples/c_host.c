/* A plain-C host of libdaftb200.so: the boundary is a C ABI (include/daft_b200.h), so any language that can call C drives the
 * kernels -- the reference's Rust host binds the same symbols (integration/daft_b200_sys.rs).
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -Ldaft_b200 -ldaftb200 -Wl,-rpath,$PWD/daft_b200 -o c_host
 *   ./c_host lineitem.parquet L_ORDERKEY L_QUANTITY
 *
 * Scans the named columns of a Parquet file into HBM (decoded on the device), hash-partitions the rows by the first column into 8
 * partitions, sorts the largest partition by its first two columns and exports the ten first rows through the Arrow C Data
 * Interface.  Every call returns a status and, on failure, a malloc'd message (released with b200_free_string). */
#include <stdio.h>
#include <stdlib.h>

#include "daft_b200.h"

#define CHECK(call)                                              \
    do {                                                         \
        char *err_ = NULL;                                       \
        if ((call) != 0) {                                       \
            fprintf(stderr, "%s\n", err_ ? err_ : "error");      \
            b200_free_string(err_);                              \
            return 1;                                            \
        }                                                        \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s file.parquet column [column...]\n", argv[0]);
        return 2;
    }
    if (b200_abi_version() != DAFT_B200_ABI_VERSION) {
        fprintf(stderr, "libdaftb200 ABI %d, header %d\n", b200_abi_version(), DAFT_B200_ABI_VERSION);
        return 2;
    }
    CHECK(b200_init(0, &err_)); /* no device -> an error: there is no CPU fallback */

    int64_t groups = 0, rows = 0;
    CHECK(b200_parquet_file_info(argv[1], &groups, &rows, &err_));
    printf("%s: %lld rows in %lld row groups\n", argv[1], (long long)rows, (long long)groups);

    b200_batch *table = NULL;
    CHECK(b200_parquet_read(argv[1], (const char *const *)(argv + 2), argc - 2, 0, -1, &table, &err_));

    const int32_t key0[1] = {0};
    b200_batch *parts[8] = {0};
    CHECK(b200_partition_by_hash(table, key0, 1, 8, parts, &err_));
    int largest = 0;
    for (int p = 1; p < 8; p++)
        if (b200_batch_num_rows(parts[p]) > b200_batch_num_rows(parts[largest])) largest = p;
    printf("partition %d holds %lld rows\n", largest, (long long)b200_batch_num_rows(parts[largest]));

    const int32_t keys[2] = {0, argc > 3 ? 1 : 0};
    const uint8_t asc[2] = {0, 0};
    b200_batch *sorted = NULL, *head = NULL;
    CHECK(b200_sort(parts[largest], keys, argc > 3 ? 2 : 1, asc, asc, &sorted, &err_));
    CHECK(b200_batch_slice(sorted, 0, 10, &head, &err_));

    struct ArrowArray array;
    struct ArrowSchema schema;
    CHECK(b200_batch_export(head, &array, &schema, &err_)); /* waits for the library stream; buffers are host memory now */
    printf("exported %lld rows x %lld columns\n", (long long)array.length, (long long)array.n_children);
    array.release(&array);
    schema.release(&schema);

    b200_batch_free(head);
    b200_batch_free(sorted);
    for (int p = 0; p < 8; p++) b200_batch_free(parts[p]);
    b200_batch_free(table);
    return 0;
}
