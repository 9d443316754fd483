/*
 * CovtGpuDecoder.java — Panama FFM (java.lang.foreign, JDK 22+) binding of libcovt_b200.so for the reference's Java tree
 * (evaluation/java/src/main/java/com/covt/decoder/). SHIPPED UN-BUILT: this image has no JDK (javac: command not found); the
 * same C symbols are exercised by the Python ctypes binding in every -m gpu test, and tests/test_abi_host.py checks that the
 * library exports every symbol bound here and that the struct layouts match what a C compiler sees. See INTEGRATION.md.
 */
package com.covt.decoder;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

/** Panama FFM binding of libcovt_b200.so: same entry point shape as CovtParser.decodeCovt, batched. No JNI-side logic. */
public final class CovtGpuDecoder implements AutoCloseable {
    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup("libcovt_b200.so", Arena.global());
    private static MethodHandle h(String name, FunctionDescriptor fd) { return LINKER.downcallHandle(LIB.find(name).orElseThrow(), fd); }

    private static final MethodHandle CREATE  = h("covt_create",  FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle DESTROY = h("covt_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LAST_ERROR = h("covt_last_error", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle DECODE_BATCH = h("covt_decode_batch",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle RESULT_NUM_LAYERS = h("covt_result_num_layers", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle RESULT_LAYERS = h("covt_result_layers", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle RESULT_TILE_STATUS = h("covt_result_tile_status", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle RESULT_BUFFER = h("covt_result_buffer", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle RESULT_READ = h("covt_result_read", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS));
    private static final MethodHandle RESULT_FREE = h("covt_result_free", FunctionDescriptor.ofVoid(ADDRESS));
    // property columns (COVT_FLAG_DECODE_PROPERTIES): CovtParser.decodePropertyColumn, CovtParser.java:276-390
    private static final MethodHandle RESULT_PROP_COLUMNS = h("covt_result_prop_columns", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle RESULT_PROP_DICTIONARIES = h("covt_result_prop_dictionaries", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle RESULT_PROP_READ = h("covt_result_prop_read", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS));
    // the library's batch scheduler: one call, several GPUs of one box (one JVM)
    private static final MethodHandle CREATE_MULTI = h("covt_create_multi", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle DECODE_BATCH_MULTI = h("covt_decode_batch_multi",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle MULTI_RESULT_PART = h("covt_multi_result_part",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle MULTI_RESULT_FREE = h("covt_multi_result_free", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle DESTROY_MULTI = h("covt_destroy_multi", FunctionDescriptor.ofVoid(ADDRESS));
    public static final long PROP_COLUMN_BYTES = 72, PROP_DICTIONARY_BYTES = 40;  // sizeof(covt_prop_column / covt_prop_dictionary)
    public static final int FLAG_DECODE_PROPERTIES = 0x80;
    // results delivered to host memory while later segments upload and decode (covt_host_sink: 13 pointers + 13 capacities in elements)
    private static final MethodHandle DECODE_BATCH_TO_HOST = h("covt_decode_batch_to_host",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    public static final StructLayout HOST_SINK = MemoryLayout.structLayout(
        MemoryLayout.sequenceLayout(13, ADDRESS).withName("ptr"), MemoryLayout.sequenceLayout(13, JAVA_LONG).withName("capacity"));
    // the stream ENCODERS: EncodingUtils.encodeVarints / encodeRle / encodeByteRle / encodeFastPfor128 (EncodingUtils.java:39-230)
    private static final MethodHandle ENCODE_STREAMS = h("covt_encode_streams",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    /** covt_encode_desc (32 bytes): value_offset (bytes), num_values, op = the covt_op that decodes the stream, num_bits | out_offset, byte_length, status. */
    public static final StructLayout ENCODE_DESC = MemoryLayout.structLayout(
        JAVA_LONG.withName("value_offset"), JAVA_INT.withName("num_values"), JAVA_BYTE.withName("op"), JAVA_BYTE.withName("num_bits"),
        MemoryLayout.paddingLayout(2), JAVA_LONG.withName("out_offset"), JAVA_INT.withName("byte_length"), JAVA_INT.withName("status"));
    public static final int OP_RLE_U64 = 3, OP_RLE_S64 = 4;
    private static final MethodHandle DECODE_STREAMS = h("covt_decode_streams",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));

    /** covt_stream_desc (48 bytes): mirrors StreamMetadata + the DecodingUtils call arguments. */
    public static final StructLayout STREAM_DESC = MemoryLayout.structLayout(
        JAVA_LONG.withName("byte_offset"), JAVA_INT.withName("byte_length"), JAVA_INT.withName("num_values"),
        JAVA_BYTE.withName("stream_type"), JAVA_BYTE.withName("encoding"), JAVA_BYTE.withName("column_type"),
        JAVA_BYTE.withName("column_data_type"), JAVA_BYTE.withName("num_bits"), JAVA_BYTE.withName("op"),
        MemoryLayout.paddingLayout(2), JAVA_INT.withName("status"), JAVA_INT.withName("bytes_consumed"),
        JAVA_LONG.withName("out_offset"), JAVA_LONG.withName("out_count"));
    public static final long LAYER_BYTES = 368;  // sizeof(covt_layer); field offsets in include/covt_b200.h

    public static final int CONTAINER_GEN2B = 0, CONTAINER_GEN3 = 1, FLAG_CLOSE_RINGS = 1;
    private final MemorySegment ctx;

    public CovtGpuDecoder(int device) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) CREATE.invokeExact(device, out);
            if (rc != 0) throw new IllegalStateException("covt_create failed (" + rc + "): " + lastError(MemorySegment.NULL));
            ctx = out.get(ADDRESS, 0);
        } catch (Throwable t) { throw sneaky(t); }
    }

    /** Batched CovtParser.decodeCovt: tiles[i] is one covtBuffer. The result owns device buffers until close(). */
    public Result decodeCovt(byte[][] tiles, int[] tileJsonFieldCounts, int container) {
        try (Arena a = Arena.ofConfined()) {
            long total = 0; for (byte[] t : tiles) total += t.length;
            MemorySegment blob = a.allocate(Math.max(total, 1), 16);
            MemorySegment offs = a.allocate(JAVA_LONG, tiles.length + 1L);
            long p = 0;
            for (int i = 0; i < tiles.length; i++) { offs.setAtIndex(JAVA_LONG, i, p); MemorySegment.copy(tiles[i], 0, blob, JAVA_BYTE, p, tiles[i].length); p += tiles[i].length; }
            offs.setAtIndex(JAVA_LONG, tiles.length, p);
            MemorySegment tj = MemorySegment.NULL;
            if (tileJsonFieldCounts != null) {           // covt_tilejson { uint32 n; const uint32* n_fields; }
                MemorySegment f = a.allocateFrom(JAVA_INT, tileJsonFieldCounts);
                tj = a.allocate(16, 8); tj.set(JAVA_INT, 0, tileJsonFieldCounts.length); tj.set(ADDRESS, 8, f);
            }
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) DECODE_BATCH.invokeExact(ctx, blob, offs, tiles.length, container, tj, FLAG_CLOSE_RINGS, out);
            if (rc != 0) throw new IllegalArgumentException("covt_decode_batch failed (" + rc + "): " + lastError(ctx));
            return new Result(out.get(ADDRESS, 0), tiles.length);
        } catch (Throwable t) { throw sneaky(t); }
    }

    /** One DecodingUtils call, e.g. decodeZigZagDeltaVarint(buf, pos, numValues): op 0 = resolve from the StreamMetadata fields. */
    public int[] decodeStream(byte[] buf, com.covt.converter.StreamType type, com.covt.converter.StreamMetadata md,
                              com.covt.converter.ColumnType columnType, int numBits, me.lemire.integercompression.IntWrapper pos) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment blob = a.allocateFrom(JAVA_BYTE, buf);
            MemorySegment d = a.allocate(STREAM_DESC);
            d.set(JAVA_LONG, 0, pos.get()); d.set(JAVA_INT, 8, md.byteLength()); d.set(JAVA_INT, 12, md.numValues());
            d.set(JAVA_BYTE, 16, (byte) type.ordinal()); d.set(JAVA_BYTE, 17, (byte) md.streamEncoding().ordinal());
            d.set(JAVA_BYTE, 18, (byte) columnType.ordinal()); d.set(JAVA_BYTE, 20, (byte) numBits);
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) DECODE_STREAMS.invokeExact(ctx, blob, (long) buf.length, d, 1, 0, out);
            if (rc != 0 || d.get(JAVA_INT, 24) != 0) throw new IllegalArgumentException("stream decode failed: " + lastError(ctx));
            MemorySegment res = out.get(ADDRESS, 0);
            int n = (int) d.get(JAVA_LONG, 40);
            MemorySegment host = a.allocate(JAVA_INT, Math.max(n, 1));
            rc = (int) RESULT_READ.invokeExact(res, 12 /*COVT_BUF_STREAM_ARENA*/, d.get(JAVA_LONG, 32), 4L * n, host);
            RESULT_FREE.invokeExact(res);
            pos.add(d.get(JAVA_INT, 28));                 // bytes_consumed: what the Java reader advanced pos by
            return host.toArray(JAVA_INT);
        } catch (Throwable t) { throw sneaky(t); }
    }

    /** EncodingUtils.encodeRle(long[] values, boolean signed) (EncodingUtils.java:123-134) on the GPU; the other encoders differ in op and value type only. */
    public byte[] encodeRle(long[] values, boolean signed) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment v = a.allocateFrom(JAVA_LONG, values);
            MemorySegment d = a.allocate(ENCODE_DESC);
            d.set(JAVA_LONG, 0, 0L); d.set(JAVA_INT, 8, values.length); d.set(JAVA_BYTE, 12, (byte) (signed ? OP_RLE_S64 : OP_RLE_U64));
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) ENCODE_STREAMS.invokeExact(ctx, v, 8L * values.length, d, 1, 0, out);
            if (rc != 0 || d.get(JAVA_INT, 28) != 0) throw new IllegalArgumentException("stream encode failed: " + lastError(ctx));
            MemorySegment res = out.get(ADDRESS, 0);
            int n = d.get(JAVA_INT, 24);
            MemorySegment host = a.allocate(JAVA_BYTE, Math.max(n, 1));
            rc = (int) RESULT_READ.invokeExact(res, 12 /*COVT_BUF_STREAM_ARENA*/, d.get(JAVA_LONG, 16), (long) n, host);
            RESULT_FREE.invokeExact(res);
            return host.asSlice(0, n).toArray(JAVA_BYTE);
        } catch (Throwable t) { throw sneaky(t); }
    }

    public final class Result implements AutoCloseable {
        final MemorySegment res; final int nTiles;
        Result(MemorySegment r, int n) { res = r; nTiles = n; }
        /** Device pointer + element count of one covt_buffer (GeoArrow-style; stays in HBM for downstream GPU consumers). */
        public long[] deviceBuffer(int which) throws Throwable {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment p = a.allocate(ADDRESS), n = a.allocate(JAVA_LONG), es = a.allocate(JAVA_INT);
                int rc = (int) RESULT_BUFFER.invokeExact(res, which, p, n, es);
                return new long[] {p.get(ADDRESS, 0).address(), n.get(JAVA_LONG, 0), es.get(JAVA_INT, 0)};
            }
        }
        /** Host copy of a slice, e.g. the coords of one layer: elemOffset = covt_layer.out[COVT_BUF_A_COORDS]. */
        public int[] readInts(int which, long elemOffset, long count) throws Throwable {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment host = a.allocate(JAVA_INT, Math.max(count, 1));
                int rc = (int) RESULT_READ.invokeExact(res, which, elemOffset, count, host);
                if (rc != 0) throw new IllegalArgumentException(lastError(ctx));
                return host.asSlice(0, 4 * count).toArray(JAVA_INT);
            }
        }
        /** The covt_prop_column records of the batch (72 bytes each, layout in include/covt_b200.h): one per property column. */
        public MemorySegment propertyColumns(int[] count) throws Throwable {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment p = a.allocate(ADDRESS), n = a.allocate(JAVA_INT);
                int rc = (int) RESULT_PROP_COLUMNS.invokeExact(res, p, n);
                if (rc != 0) throw new IllegalArgumentException(lastError(ctx));
                count[0] = n.get(JAVA_INT, 0);
                return p.get(ADDRESS, 0).reinterpret(PROP_COLUMN_BYTES * count[0]);
            }
        }
        /** INT_64 property column -> Optional-like long[] + present[] (CovtParser.java:296-326): slot i belongs to feature i. */
        public long[] readLongColumn(MemorySegment col, boolean[] present) throws Throwable {
            try (Arena a = Arena.ofConfined()) {
                int F = col.get(JAVA_INT, 40);                                   // num_features
                long validityOffset = col.get(JAVA_LONG, 48), valuesOffset = col.get(JAVA_LONG, 56);
                MemorySegment bits = a.allocate(Math.max((F + 7) / 8, 1)), vals = a.allocate(JAVA_LONG, Math.max(F, 1));
                int rc = (int) RESULT_PROP_READ.invokeExact(res, 0 /*COVT_PBUF_VALIDITY*/, validityOffset, (long) ((F + 7) / 8), bits);
                if (rc == 0) rc = (int) RESULT_PROP_READ.invokeExact(res, 1 /*COVT_PBUF_I64*/, valuesOffset, (long) F, vals);
                if (rc != 0) throw new IllegalArgumentException(lastError(ctx));
                for (int i = 0; i < F; i++) present[i] = ((bits.get(JAVA_BYTE, i >> 3) >> (i & 7)) & 1) != 0;
                return vals.asSlice(0, 8L * F).toArray(JAVA_LONG);
            }
        }
        @Override public void close() { try { RESULT_FREE.invokeExact(res); } catch (Throwable t) { throw sneaky(t); } }
    }

    private static String lastError(MemorySegment c) {
        try (Arena a = Arena.ofConfined()) { MemorySegment b = a.allocate(512); int rc = (int) LAST_ERROR.invokeExact(c, b, 512L); return b.getString(0); }
        catch (Throwable t) { return "?"; }
    }
    @Override public void close() { try { DESTROY.invokeExact(ctx); } catch (Throwable t) { throw sneaky(t); } }
    private static RuntimeException sneaky(Throwable t) { return t instanceof RuntimeException r ? r : new RuntimeException(t); }
}
