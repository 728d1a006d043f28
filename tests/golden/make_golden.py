"""Regenerates tests/golden/test_query.{paf,sam} with the CPU oracle from the reference's own tiny fixtures
(copied from /root/reference/data: test_ref.fasta, test_ref.gtf, test_query.fastq, flags of data/Makefile:21).
The reference binary cannot be built in this image, so these are ORACLE outputs (they agree with the table in
SURVEY.md section 4 that was derived independently); rerun when a box with cargo can produce the real thing."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import orc  # noqa: E402

rd = lambda n: open(os.path.join(HERE, n), "rb").read()
ix = orc.Index.create(rd("test_ref.fasta"), rd("test_ref.gtf"))
fq = rd("test_query.fastq")
open(os.path.join(HERE, "test_query.paf"), "wb").write(ix.align_fastq_text(fq, k=3, min_score=0))
open(os.path.join(HERE, "test_query.sam"), "wb").write(ix.align_fastq_text(fq, k=3, min_score=0, sam=True))
