#!/bin/sh
# Same command line as the reference's hw5/run.sh: ./run.sh <scene.txt> <out.ppm>
exec "$(dirname "$0")/raytracing-course_b200/raytracing_hw5" "$1" "$2"
