/* Empty stand-in. The reference snapshot lists this header in .MISSING_LARGE_BLOBS;
   bsdfmodel/holzschuchpacanowski.h includes it but no model reads its tables. */
