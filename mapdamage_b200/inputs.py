"""Host glue: alignment file + FASTA -> (ReadBatch, Reference, libraries)."""
from .batch import BatchBuilder
from .refgenome import Reference
from .samtext import read_sam


def load_alignments(sam_path, fasta_path, merge_libraries=False, apply_filter=True):
    """Reads a SAM text file into one batch.

    Returns ``(batch, reference, libraries, header)`` with the reference's
    contigs in BAM-header (tid) order and ``libraries`` the sorted
    ``(sample, library)`` list that indexes the count slabs.
    """
    header, records = read_sam(sam_path)
    reference = Reference.from_fasta(fasta_path).reordered(header.references, header.lengths)
    builder = BatchBuilder(
        readgroups=None if merge_libraries else header.libraries(),
        merge_libraries=merge_libraries, apply_filter=apply_filter,
    )
    for record in records:
        builder.add(record)
    return builder.finish(), reference, builder.libraries, header
