"""Helper namespace mirroring ``vndecorrelate.utils`` for the hot path (see ``dsp``)."""
