// DOT handler fragment of vm_product (measured on B200: -10 % against the Karatsuba programs, see profiles/experiments_r2.txt)
        if (BNP_DOT_OP && op == BNP_OP_DOT) {
            // further terms: the three Karatsuba parts of every term are added to the running wide sums
            // A = sum x0 y0, B = sum x1 y1, C = sum (x0 + x1)(y0 + y1) (at most 4 terms: C < 16 p^2 < 2^512), and the
            // combination below plus ONE pair of reductions finishes the whole dot product
            const u32 nt = (imm >> BNP_DOT_N_SHIFT) & 3u;
            u64 tw = (u64)(b | (ee << 8)) | ((u64)dotw << 16);
#pragma unroll 1
            for (u32 t = 0; t < nt; t++) {
                S.load(x, (u32)tw & 0xffu);
                S.load(y, ((u32)tw >> 8) & 0xffu);
                tw >>= 16;
                u32 Q[16];
                add8(sx, x.c0, x.c1);
                add8(sy, y.c0, y.c1);
                fp_mul_wide(Q, sx, sy);
                add16(P2, Q);
                fp_mul_wide(Q, x.c0, y.c0);
                add16(T0, Q);
                fp_mul_wide(Q, x.c1, y.c1);
                add16(T1, Q);
            }
        }
