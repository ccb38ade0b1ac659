"""Path constants of the host mirror (reference: src/config.py:23-29).

The reference resolves 'embedders' / 'models' relative to the CWD; here the defaults are absolute paths inside the
package so the plugins load from anywhere.  Point BUZZ_B200_PLUGIN_ROOT at a buzzdetect checkout whose plugin files
were replaced by the ones in buzzdetect_b200/plugins/ to use that tree instead."""
import os

_ROOT = os.environ.get("BUZZ_B200_PLUGIN_ROOT",
                       os.path.join(os.path.dirname(os.path.abspath(__file__)), "plugins"))

DIR_EMBEDDERS = os.path.join(_ROOT, "embedders")
DIR_MODELS = os.path.join(_ROOT, "models")
DEFAULT_MODEL = "model_general_v3"
SUBDIR_TESTS = "tests"
FNAME_METRICS = "metrics.csv"

SUFFIX_RESULT_COMPLETE = "_buzzdetect.csv"
SUFFIX_RESULT_PARTIAL = "_buzzpart.csv"
PREFIX_COLUMN_ACTIVATION = "activation_"
PREFIX_COLUMN_DETECTION = "detections_"
