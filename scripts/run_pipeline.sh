#!/bin/bash
# run_pipeline.sh -- BASELINE.json config 5: the reference's run.sh sequence around the match stage, with the stages
# this repo does not rebuild replaced by stand-ins (surf3d and frog need VTK/OpenCV, absent here).
#
#   run_pipeline.sh <params.sh> <keypoint-dir> <match-executable> [frog-stand-in]
#
# Follows /root/reference/run.sh step by step (cited by line):
#   :29      source the parameter file (RES_FOLDER, NPOINTS, MAX_DISTANCE, MATCH_OTHER_PARAMS, IMG_INPUT[])
#   :57-70   create the result folder           :72-78  cd into it, start an empty points.txt
#   :81-88   per image: "surf3d IMG -o $RES_FOLDER/points$k ..." then append ${OUTPUT_POINTS}.csv.gz to points.txt
#            -- stand-in: the image's pre-generated keypoint file <keypoint-dir>/points$k.csv.gz is copied into place
#   :92-93   "$MATCH points.txt -o pairs.bin -d $MAX_DISTANCE $MATCH_OTHER_PARAMS"   <- the stage under test, verbatim
#   :104     "$REG pairs.bin ..."  -- stand-in: a reader of pairs.bin restating ImageGroup::readPairs
#            (registration/imageGroup.cxx:1353-1417), which aborts on a malformed file
#   :108-111 per-stage wall seconds
startTime=`date +%s%N`
. $1
KP_DIR=$2
MATCH=$3
REG=${4:-true}
IMG_NUMBER=${#IMG_INPUT[@]}
if [ ${IMG_NUMBER} -lt 2 ]; then echo "Less than 2 images have been specified"; exit 1; fi
if [ ! -d ${RES_FOLDER} ]; then mkdir -p ${RES_FOLDER} || exit 1; fi
cd $RES_FOLDER
PointsFile=points.txt
if [ -f ${PointsFile} ]; then rm ${PointsFile}; fi
for (( CUR_IT=0; CUR_IT<${IMG_NUMBER}; CUR_IT++ )); do
  OUTPUT_POINTS=$RES_FOLDER/points$CUR_IT
  cp $KP_DIR/points$CUR_IT.csv.gz ${OUTPUT_POINTS}.csv.gz || exit 1   # stand-in for: $SURF $IMG -o $OUTPUT_POINTS -s $SPACING -t $THRESHOLD -n $NPOINTS
  echo ${OUTPUT_POINTS}.csv.gz>> ${PointsFile}
done
matchTime=`date +%s%N`
OUTPUT_PAIRS=pairs.bin
echo "Executing : $MATCH $PointsFile -o $OUTPUT_PAIRS -d $MAX_DISTANCE $MATCH_OTHER_PARAMS"
$MATCH $PointsFile -o $OUTPUT_PAIRS -d $MAX_DISTANCE $MATCH_OTHER_PARAMS || exit 1
registrationTime=`date +%s%N`
$REG $OUTPUT_PAIRS $REGISTRATION_OTHER_PARAMS || exit 1
endTime=`date +%s%N`
echo "Keypoint extraction time : $(( (matchTime-startTime)/1000000 )) ms"
echo "Match time : $(( (registrationTime-matchTime)/1000000 )) ms"
echo "Registration time : $(( (endTime-registrationTime)/1000000 )) ms"
