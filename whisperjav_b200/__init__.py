"""whisperjav_b200 -- B200-native backend for WhisperJAV's ASR hot path.

``register()`` plugs the backend into the reference's string-keyed factories when WhisperJAV is
importable (module-level dict mutation is the only registration mechanism it has):

    import whisperjav_b200; whisperjav_b200.register()
    from whisperjav.main import main; main()     # e.g. --pipeline decoupled --generator b200-whisper
"""
__version__ = "0.1.0"


def register() -> dict:
    """Returns {surface: registered?}.  Safe to call when WhisperJAV is not installed."""
    done = {"speech_segmenter": False, "text_generator": False, "scene_detector": False}
    try:
        from whisperjav.modules.speech_segmentation import factory as sf  # type: ignore
        sf._BACKEND_REGISTRY["b200-vad"] = "whisperjav_b200.segmenter.B200SpeechSegmenter"
        sf._BACKEND_DEPENDENCIES["b200-vad"] = {"packages": [], "install_hint": "", "always_available": True}
        sf._BACKEND_REGISTRY["b200-whisperseg"] = "whisperjav_b200.whisperseg.B200WhisperSegSegmenter"
        sf._BACKEND_DEPENDENCIES["b200-whisperseg"] = {"packages": [], "install_hint": "", "always_available": True}
        done["speech_segmenter"] = True
    except Exception:
        pass
    try:
        from whisperjav.modules.subtitle_pipeline.generators import factory as gf  # type: ignore
        gf._REGISTRY["b200-whisper"] = "whisperjav_b200.generator.B200WhisperGenerator"
        done["text_generator"] = True
    except Exception:
        pass
    try:
        from whisperjav.modules.scene_detection_backends import factory as scf  # type: ignore
        scf._BACKEND_REGISTRY["b200-auditok"] = "whisperjav_b200.scenes.B200SceneDetector"
        scf._BACKEND_DEPENDENCIES["b200-auditok"] = {"packages": [], "install_hint": "", "always_available": True}
        scf._BACKEND_REGISTRY["b200-silero"] = "whisperjav_b200.scenes.B200SileroSceneDetector"
        scf._BACKEND_DEPENDENCIES["b200-silero"] = {"packages": [], "install_hint": "", "always_available": True}
        done["scene_detector"] = True
    except Exception:
        pass
    return done
