"""CPU oracle for the EveryVoice preprocessing feature-extraction hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.
``everyvoice_b200`` (the product) never imports anything from here and has no CPU
fallback: without ``libevfeat.so`` and a CUDA device it raises.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the *live* reference from
``/root/reference`` (in the build container only), runs the reference's own functions
(``everyvoice.utils.heavy.get_spectral_transform``, ``Preprocessor.extract_*``,
``Preprocessor.average_data_by_durations``, ``helpers.Scaler``) on seeded inputs and
writes their outputs to ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
this restatement against those vectors on every CPU test run.  The one piece that cannot
be pinned against live reference code is ``librosa.filters.mel`` (librosa 0.11.0 is not
installed here); it is restated from its published algorithm and cross-checked against
``torchaudio.functional.melscale_fbanks(mel_scale="slaney", norm="slaney")``.
"""
